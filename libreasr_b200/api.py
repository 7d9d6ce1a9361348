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
        self.denumericalize = denumericalize or (lambda ids: list(ids))

    @property
    def engine(self):
        """Fetched from the model on every use: `m.lm = ...`, `load_state_dict` or `.to()` rebuild the engine, a cached
        reference would point at a destroyed handle."""
        return self.model.engine()

    # offline: ASRServicer.Transcribe (api-server.py:64-80) for one or many utterances
    def transcribe(self, audio, lens=None, max_iters=3, sr=16000):
        a = _as_audio(audio)
        single = a.shape[0] == 1 and not (torch.is_tensor(audio) and audio.dim() == 2)
        if int(sr) != self.engine.cfg.sample_rate:   # Resample (transforms.py:135-144) on the device, then the 16 kHz path
            if lens is not None:
                raise ValueError("lens are not supported together with resampling; trim the utterances first")
            a = self.engine.resample(a.to(self.engine.device), int(sr))
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
        sb = None
        for ch in chunks:
            a = _as_audio(ch)
            if sb is None:  # the first chunk fixes the chunk size of the stream (80 ms in the reference client)
                sb = StreamBatch(self.engine, 1, chunk=a.shape[1], max_iters=max_iters)
            new = sb.push(a)
            if new is not None:
                yield list(sb.tokens[0]), self.denumericalize(new[0])


class StreamBatch:
    """B concurrent streams, one 80 ms chunk per (active) stream per ``push``; every stream runs its own phase of the
    serving loop and can be reset (new connection) or skipped (no chunk this tick) independently.

    Reproduces, per stream, the reference serving loop: 3-chunk sliding window
    (api-server.py:26,95-102) -> stream transforms incl. ``Buffer(n_buffer=2)``
    (config/testing.yaml:356-374) -> ``Transducer.transcribe_stream`` with carried encoder /
    predictor / LM-fuser state (models.py:457-577).  A thin veneer over the C ABI's streaming session
    (``rnnt_b200_stream_open/push/reset/close``): one library call per chunk tick; all state (audio window,
    pending feature rows, LSTM (h, c), GRU h, last predictor output, LM fuser) stays resident in HBM."""

    def __init__(self, engine: Engine, n_streams: int, chunk: int = None, max_iters: int = 10, n_buffer: int = 2):
        import ctypes as C

        cfg = engine.cfg
        self.engine, self.B, self.max_iters, self.n_buffer = engine, n_streams, max_iters, n_buffer
        self.chunk = chunk or cfg.sample_rate * CHUNK_MS // 1000
        self._s = C.c_void_p(0)
        with torch.cuda.device(engine.device):
            engine._ck(engine.lib.rnnt_b200_stream_open(engine._h, n_streams, self.chunk, BUFFER_N_FRAMES, n_buffer, max_iters,
                                                        C.byref(self._s)))
        engine._register_session(self)
        self.U = max_iters * n_buffer
        self._tok = torch.zeros(n_streams, self.U, dtype=torch.int32).pin_memory()
        self._ntok = torch.zeros(n_streams, dtype=torch.int32).pin_memory()
        self._adv = C.c_int32(0)
        self.tokens = [[] for _ in range(n_streams)]

    def close(self):
        if getattr(self, "_s", None) is not None and self._s.value:
            self.engine.lib.rnnt_b200_stream_close(self._s)
            self._s.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, slot=None):
        """``slot`` (or every stream when None) back to the state of a fresh connection (models.py:494-500)."""
        self.engine._ck(self.engine.lib.rnnt_b200_stream_reset(self._s, -1 if slot is None else int(slot)))
        for b in (range(self.B) if slot is None else [int(slot)]):
            self.tokens[b] = []

    def reset_state(self, slot=None):
        """Mid-stream reset = the ``reset_fn`` of ``transcribe_stream`` (models.py:480-500; the server calls it on
        silence, api-server.py:133-135): encoder / predictor / LM-fuser state only -- the audio window and the Buffer
        of the serving loop keep their contents (unlike ``reset``, which is a new connection)."""
        self.engine._ck(self.engine.lib.rnnt_b200_stream_reset_state(self._s, -1 if slot is None else int(slot)))

    def push(self, chunks, active=None):
        """chunks [B, chunk] float32, CUDA or (ideally pinned) CPU tensor; ``active``: optional [B] booleans, streams with
        False are skipped this tick.  Returns the list of new token lists (empty for streams that did not run) when at
        least one stream advanced its encoder on this call, else None."""
        import ctypes as C

        eng = self.engine
        if chunks.dtype != torch.float32 or not chunks.is_contiguous():
            chunks = chunks.to(torch.float32).contiguous()
        if tuple(chunks.shape) != (self.B, self.chunk):
            raise ValueError(f"expected chunks of shape {(self.B, self.chunk)}, got {tuple(chunks.shape)}")
        with torch.cuda.device(eng.device):
            act = None
            if active is not None:
                act = np.ascontiguousarray(np.asarray(active, dtype=np.uint8))
                if act.shape != (self.B,):
                    raise ValueError(f"expected {self.B} activity flags")
            eng._ck(eng.lib.rnnt_b200_stream_push(self._s, C.c_void_p(chunks.data_ptr()), 0 if chunks.is_cuda else 1,
                                                  C.c_void_p(act.ctypes.data) if act is not None else C.c_void_p(0),
                                                  C.c_void_p(self._tok.data_ptr()), self.U, C.c_void_p(self._ntok.data_ptr()),
                                                  C.byref(self._adv), eng._stream()))
        if not self._adv.value:
            return None
        t, n = self._tok.numpy(), self._ntok.numpy()
        self.ran = [int(n[b]) >= 0 for b in range(self.B)]   # which streams took part in this model step
        new = [t[b, : max(int(n[b]), 0)].tolist() for b in range(self.B)]
        for b in range(self.B):
            self.tokens[b].extend(new[b])
        return new
