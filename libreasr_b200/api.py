"""``LibreASR`` facade (the ``LibreASR.transcribe()/stream()`` surface BASELINE.json names;
at the reference snapshot the same role is played by ``load_stuff`` + ``ASRServicer``,
libreasr/lib/inference.py:18-51, api-server.py:53-135) and the batched multi-stream
scheduler used for 64 concurrent 80 ms streams."""
import numpy as np
import torch

from .engine import Engine, EngineConfig, tokens_to_lists

CHUNK_MS = 80  # api-client.py:14
BUFFER_N_FRAMES = 3  # api-server.py:26


def _as_audio(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(a, dtype=np.float32).copy()  # tensorize, utils.py:149-153
    t = torch.as_tensor(a, dtype=torch.float32)
    return t[None] if t.dim() == 1 else t


class LibreASR:
    """``LibreASR(model).transcribe(audio)`` / ``.stream(chunks)``.

    ``model`` is a ``libreasr_b200.lib.models.Transducer`` (weights loaded, on a CUDA
    device).  ``denumericalize`` maps token ids to text (the reference uses a
    youtokentome BPE model, language.py:115-151); by default ids are returned."""

    def __init__(self, model, denumericalize=None):
        self.model = model
        self.engine = model.engine()
        self.denumericalize = denumericalize or (lambda ids: list(ids))

    # offline: ASRServicer.Transcribe (api-server.py:64-80) for one or many utterances
    def transcribe(self, audio, lens=None, max_iters=3, sr=16000):
        if int(sr) != self.engine.cfg.sample_rate:
            raise NotImplementedError("only 16 kHz input is built (Resample is an identity on this path)")
        a = _as_audio(audio)
        single = a.shape[0] == 1 and not (torch.is_tensor(audio) and audio.dim() == 2)
        if a.is_cuda:
            r = self.engine.transcribe(a, lens, max_iters)
        else:
            a = a.contiguous()
            r = self.engine.transcribe_host(a, None if lens is None else torch.as_tensor(lens, dtype=torch.int32), max_iters)
        outs = [self.denumericalize(t) for t in tokens_to_lists(r["tokens"], r["ntok"])]
        return outs[0] if single else outs

    # streaming: ASRServicer.TranscribeStream (api-server.py:82-135) for ONE stream
    def stream(self, chunks, max_iters=10):
        """``chunks``: iterable of 80 ms float32 PCM chunks (arrays or raw bytes).  Yields
        (all token ids so far, denumericalized tokens of this step) each time the model
        advanced (every second chunk once three are buffered)."""
        sb = StreamBatch(self.engine, 1, max_iters=max_iters)
        for ch in chunks:
            a = _as_audio(ch)
            new = sb.push(a.to(self.engine.device))
            if new is not None:
                yield list(sb.tokens[0]), self.denumericalize(new[0])


class StreamBatch:
    """B concurrent streams advanced in lock step, one 80 ms chunk per stream per ``push``.

    Reproduces, per stream, the reference serving loop: 3-chunk sliding window
    (api-server.py:26,95-102) -> stream transforms incl. ``Buffer(n_buffer=2)``
    (config/testing.yaml:356-374) -> ``Transducer.transcribe_stream`` with carried encoder /
    predictor state (models.py:457-577).  All state (audio window, pending feature row,
    LSTM (h, c), GRU h, last predictor output) stays resident in HBM."""

    def __init__(self, engine: Engine, n_streams: int, chunk: int = None, max_iters: int = 10, n_buffer: int = 2):
        cfg = engine.cfg
        self.engine, self.B, self.max_iters, self.n_buffer = engine, n_streams, max_iters, n_buffer
        self.chunk = chunk or cfg.sample_rate * CHUNK_MS // 1000
        dev = engine.device
        self.window = torch.zeros(n_streams, BUFFER_N_FRAMES * self.chunk, device=dev)
        self.n_chunks = 0
        self.rows = []
        self.enc_state = None
        self.pred_state = None
        # one LM fuser per stream (models.py:478), resident on the device like the other stream state
        self.lm_state = engine.new_lm_state(n_streams) if cfg.lm_layers > 0 else None
        self.tokens = [[] for _ in range(n_streams)]

    def reset(self):
        self.enc_state, self.pred_state = None, None
        if self.lm_state is not None:
            self.lm_state.zero_()  # LMFuser.reset (lm.py:81-83)

    def push(self, chunks):
        """chunks [B, chunk] CUDA tensor.  Returns the list of new token lists when the
        encoder advanced on this call, else None."""
        c = self.chunk
        # slide by one chunk (api-server.py:99-102)
        self.window = torch.cat([self.window[:, c:], chunks.to(self.window.dtype)], dim=1)
        self.n_chunks += 1
        if self.n_chunks < BUFFER_N_FRAMES:
            return None
        self.rows.append(self.engine.features_stream(self.window))  # [B, X]
        if len(self.rows) < self.n_buffer:
            return None
        feats = torch.stack(self.rows, dim=1)  # [B, n_buffer, X]
        self.rows.clear()
        enc, self.enc_state = self.engine.encode(feats, state=self.enc_state, want_state=True)
        r = self.engine.decode_greedy(enc, max_iters=self.max_iters, state=self.pred_state, want_state=True, lm_state=self.lm_state)
        self.pred_state = r["state"]
        new = tokens_to_lists(r["tokens"], r["ntok"])
        for b in range(self.B):
            self.tokens[b].extend(new[b])
        return new
