"""Serving wire path around the streaming session ("next" row 2).

What the reference does per connection in Python threads (``ASRServicer``, api-server.py:53-135: one gRPC worker thread per
stream, one shared model, one shared ``Buffer``) is rebuilt as ONE scheduler thread that advances all connected streams
together through ``rnnt_b200_stream_push`` (``StreamBatch``): every connection owns a slot of the session, its frames are
queued, and each scheduler tick takes at most one pending 80 ms frame per slot, pushes the batch and routes the new
tokens back.  The wire formats are the reference's:

* gRPC service ``ASR.ASR`` with ``Transcribe(Audio) -> Transcript`` and ``TranscribeStream(stream Audio) -> stream
  Transcript`` (interfaces/libreasr.proto:5-17); ``Audio{bytes data = 1; int32 sr = 3}``, ``Transcript{string data = 1}``.
  No generated stubs are needed: the two tiny messages are (de)serialised here in protobuf wire format and the
  service is registered through gRPC's generic handlers.
* the WebSocket bridge frame ``[4 B ascii lang][f32 sample rate][f32 PCM ...]`` (api-bridge.py:96-106).
* the transcript post-processing of ``TranscribeStream`` (api-server.py:117-135): emit the characters that changed,
  drop a repeated diff, and reset the stream state after ~4 s when a step produced nothing new.
"""
import itertools as it
import queue
import struct
import threading
from concurrent import futures

import numpy as np
import torch

THRESH = 4000        # api-server.py:22
SERVICE = "ASR.ASR"  # package ASR; service ASR (libreasr.proto:3-5)


# ---------------------------------------------------------------------------------------------------------------
# protobuf wire format of the two messages (libreasr.proto:10-17)
# ---------------------------------------------------------------------------------------------------------------
def _varint(n: int) -> bytes:
    n &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        out.append(b | (0x80 if n else 0))
        if not n:
            return bytes(out)


def _read_varint(b: bytes, i: int):
    shift = v = 0
    while True:
        if i >= len(b):
            raise ValueError("truncated varint")
        c = b[i]
        i += 1
        v |= (c & 0x7F) << shift
        if not c & 0x80:
            return v, i
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _fields(b: bytes):
    """Yields (field number, wire type, value) of a serialized message; unknown fields are returned too."""
    i = 0
    while i < len(b):
        key, i = _read_varint(b, i)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, i = _read_varint(b, i)
        elif wt == 2:
            n, i = _read_varint(b, i)
            if i + n > len(b):
                raise ValueError("truncated length-delimited field")
            v, i = b[i:i + n], i + n
        elif wt == 1:
            v, i = b[i:i + 8], i + 8
        elif wt == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def encode_audio(data: bytes, sr: int) -> bytes:
    """``Audio{bytes data = 1; int32 sr = 3}`` (proto3: default values are not written)."""
    out = b""
    if data:
        out += b"\x0a" + _varint(len(data)) + bytes(data)
    if sr:
        out += b"\x18" + _varint(int(sr))
    return out


def decode_audio(b: bytes):
    data, sr = b"", 0
    for fno, wt, v in _fields(b):
        if fno == 1 and wt == 2:
            data = bytes(v)
        elif fno == 3 and wt == 0:   # int32 on the wire: sign-extended to 64 bits, keep the low 32 as a signed value
            sr = (v & 0xFFFFFFFF) - (1 << 32) if v & 0x80000000 else v & 0xFFFFFFFF
    return data, sr


def encode_transcript(text: str) -> bytes:
    raw = text.encode("utf-8")
    return (b"\x0a" + _varint(len(raw)) + raw) if raw else b""


def decode_transcript(b: bytes) -> str:
    text = ""
    for fno, wt, v in _fields(b):
        if fno == 1 and wt == 2:
            text = bytes(v).decode("utf-8")
    return text


# ---------------------------------------------------------------------------------------------------------------
# WebSocket bridge frame (api-bridge.py:96-106)
# ---------------------------------------------------------------------------------------------------------------
def parse_ws_frame(payload: bytes):
    """``[4 B ascii language, space padded][float32 sample rate][float32 PCM ...]`` -> (lang, sr, pcm bytes)."""
    if len(payload) < 8:
        raise ValueError("frame shorter than its 8-byte header")
    lang = payload[:4].decode("ascii").strip()
    sr = int(struct.unpack("f", payload[4:8])[0])
    return lang, sr, payload[8:]


def build_ws_frame(lang: str, sr: int, pcm) -> bytes:
    """What the web / ESP32 clients send (apps/web/src/lib/utils.js:20-44)."""
    head = lang.encode("ascii")[:4].ljust(4, b" ") + struct.pack("f", float(sr))
    return head + np.asarray(pcm, dtype=np.float32).tobytes()


def tensorize(data: bytes) -> np.ndarray:
    """utils.py:149-153: raw little-endian float32 bytes -> [1, n]."""
    return np.frombuffer(data, dtype="<f4").astype(np.float32)[None]


# ---------------------------------------------------------------------------------------------------------------
# transcript post-processing of TranscribeStream (api-server.py:44-50, 117-135)
# ---------------------------------------------------------------------------------------------------------------
def should_reset(steps: int, downsample: int, n_buffer: int) -> bool:
    return int(10.0 * downsample * n_buffer * steps) >= THRESH


class TranscriptDiffer:
    """One per connection.  ``step`` takes the output of one model step -- all token ids so far and the ids of this step --
    and returns (text to send or None, reset the stream?)."""

    def __init__(self, denumericalize, downsample=8, n_buffer=2):
        self.denumericalize, self.downsample, self.n_buffer = denumericalize, downsample, n_buffer
        self.last, self.last_diff, self.steps = "", "", 0

    def step(self, y_all, y_new):
        self.steps += 1
        if self.denumericalize(y_new) != "":
            now = self.denumericalize(y_all)
            diff = "".join(y for x, y in it.zip_longest(self.last, now) if x != y)
            self.last = now
            if diff == self.last_diff:   # "bail if we just output the same thing twice"
                return None, False
            self.last_diff = diff
            return diff, False
        if should_reset(self.steps, self.downsample, self.n_buffer):
            self.steps = 0
            return None, True
        return None, False


# ---------------------------------------------------------------------------------------------------------------
# scheduler: connections <-> slots of one streaming session
# ---------------------------------------------------------------------------------------------------------------
_CLOSE = object()   # in-band end-of-stream marker of a connection's frame queue


class StreamScheduler:
    """``session`` exposes ``B``, ``chunk``, ``push(chunks [B, chunk] float32 tensor, active=[B] bools) -> list of token lists
    | None`` and ``reset(slot)`` -- i.e. ``libreasr_b200.api.StreamBatch``.  All session calls happen on the thread that calls
    ``tick()`` / ``run()``; ``connect`` / ``feed`` / ``disconnect`` may be called from any thread."""

    def __init__(self, session, lock=None):
        self.session, self.B, self.chunk = session, session.B, session.chunk
        self.lock = lock or threading.Lock()          # serialises use of the engine with other callers (unary Transcribe)
        self._mu = threading.Lock()
        self._free = list(range(self.B - 1, -1, -1))
        self._inq = [queue.Queue() for _ in range(self.B)]
        self._outq = [None] * self.B
        self._live = [False] * self.B
        self._fresh = [False] * self.B                # connected since the last tick: reset the slot before its first frame
        self._all = [[] for _ in range(self.B)]
        self._wake = threading.Event()
        self._buf = torch.zeros(self.B, self.chunk, dtype=torch.float32)
        try:
            self._buf = self._buf.pin_memory()
        except Exception:
            pass  # no CUDA runtime (unit tests with a fake session)
        self.ticks = self.model_ticks = 0

    # ---- connection side ----
    def connect(self):
        with self._mu:
            if not self._free:
                raise RuntimeError(f"all {self.B} stream slots are in use")
            slot = self._free.pop()
            self._inq[slot] = queue.Queue()
            self._outq[slot] = queue.Queue()
            self._live[slot], self._fresh[slot], self._all[slot] = True, True, []
            return slot

    def feed(self, slot, frame):
        """One frame of ``chunk`` float32 samples (the reference client sends 80 ms, api-client.py:14)."""
        a = np.asarray(frame, dtype=np.float32).reshape(-1)
        if a.size != self.chunk:
            raise ValueError(f"streaming frames must hold {self.chunk} samples, got {a.size}")
        self._inq[slot].put(a)
        self._wake.set()

    def request_reset(self, slot):
        self._inq[slot].put(None)   # takes effect in order with the frames
        self._wake.set()

    def close(self, slot):
        """End of the client's stream: handled by the scheduler thread in order, i.e. after every frame fed before it has
        been pushed and its tokens delivered; then the end-of-stream marker is queued and the slot is freed."""
        self._inq[slot].put(_CLOSE)
        self._wake.set()

    def results(self, slot):
        return self._outq[slot]

    def disconnect(self, slot):
        with self._mu:
            if self._live[slot]:
                self._live[slot] = False
                self._outq[slot].put(None)          # end-of-stream marker for the consumer
                self._free.append(slot)

    # ---- scheduler side ----
    def pending(self):
        return any(self._live[b] and not self._inq[b].empty() for b in range(self.B))

    def tick(self):
        """Takes at most one pending frame per live slot and advances those streams by one chunk."""
        active = [False] * self.B
        with self._mu:   # one consistent snapshot: which slots are live, which are new connections, and THEIR queues
            live = list(self._live)
            fresh, self._fresh = self._fresh, [False] * self.B
            inq, outq = list(self._inq), list(self._outq)
        resets = [b for b in range(self.B) if live[b] and fresh[b]]   # new connections: full reset (window, Buffer, state)
        state_resets = []                                              # reset_fn requests: model state only (models.py:480-500)
        closing = []
        for b in range(self.B):
            if not live[b]:
                continue
            while True:
                try:
                    item = inq[b].get_nowait()
                except queue.Empty:
                    break
                if item is None:                      # reset request from the transcript logic
                    if b not in resets and b not in state_resets:
                        state_resets.append(b)
                    continue
                if item is _CLOSE:                    # everything fed before it has been pushed in earlier ticks
                    closing.append(b)
                    break
                self._buf[b].copy_(torch.from_numpy(item))
                active[b] = True
                break
        for b in closing:
            self.disconnect(b)
        if not resets and not state_resets and not any(active):
            return bool(closing)
        with self.lock:
            for b in resets:
                self.session.reset(b)
            for b in state_resets:
                getattr(self.session, "reset_state", self.session.reset)(b)
            new = self.session.push(self._buf, active=active) if any(active) else None
        self.ticks += 1
        if new is not None:
            self.model_ticks += 1
            for b in range(self.B):
                # a stream takes part in a model step with >= 0 tokens when it was active this tick and its Buffer filled;
                # the session reports [] for streams that did not run, which the transcript logic must not count as a step
                if active[b] and live[b] and self._ran(b):
                    self._all[b] = self._all[b] + list(new[b])
                    outq[b].put((list(self._all[b]), list(new[b])))
        return True

    def _ran(self, b):
        ran = getattr(self.session, "ran", None)      # optional per-stream "took part in the model step" flags
        return True if ran is None else bool(ran[b])

    def run(self, stop: threading.Event, idle_s=0.002):
        """Scheduler loop.  A failing tick must not leave the connected clients blocked on their result queues: every
        live slot is disconnected (end-of-stream marker) and the error is kept in ``self.error`` and re-raised."""
        try:
            while not stop.is_set():
                if not self.tick():
                    self._wake.wait(idle_s)
                    self._wake.clear()
        except Exception as e:  # noqa: BLE001
            self.error = e
            for b in range(self.B):
                self.disconnect(b)
            raise


# ---------------------------------------------------------------------------------------------------------------
# gRPC service (api-server.py:53-152) through generic handlers
# ---------------------------------------------------------------------------------------------------------------
class ASRServicer:
    def __init__(self, asr, scheduler=None, n_streams=64, denumericalize=None, downsample=8, n_buffer=2):
        """``asr``: a ``libreasr_b200.LibreASR``; ``denumericalize``: token ids -> str (default: the facade's)."""
        from .api import StreamBatch

        self.asr = asr
        self.denum = denumericalize or (lambda ids: "".join(chr(0x100 + t) for t in ids))
        self.lock = threading.Lock()
        self.scheduler = scheduler or StreamScheduler(StreamBatch(asr.engine, n_streams, n_buffer=n_buffer), lock=self.lock)
        self.downsample, self.n_buffer = downsample, n_buffer
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self.scheduler.run, args=(self._stop,), daemon=True)
        self._thread.start()

    def close(self):
        self._stop.set()
        self._thread.join(timeout=5)

    # rpc Transcribe(Audio) returns (Transcript)
    def Transcribe(self, request: bytes, context):
        data, sr = decode_audio(request)
        aud = tensorize(data)
        with self.lock:
            ids = self.asr.transcribe(torch.from_numpy(aud[0].copy()), sr=sr or self.asr.engine.cfg.sample_rate)
        return encode_transcript(ids if isinstance(ids, str) else self.denum(ids))

    # rpc TranscribeStream(stream Audio) returns (stream Transcript)
    def TranscribeStream(self, request_iterator, context):
        import grpc

        sch = self.scheduler
        try:
            slot = sch.connect()
        except RuntimeError as e:
            context.abort(grpc.StatusCode.RESOURCE_EXHAUSTED, str(e))
        differ = TranscriptDiffer(self.denum, self.downsample, self.n_buffer)
        err = []

        def reader():
            try:
                for req in request_iterator:
                    data, sr = decode_audio(req)
                    pcm = tensorize(data)[0]
                    if sr and sr != self.asr.engine.cfg.sample_rate:
                        # per-frame Resample like the reference's stream pipeline (transforms.py:135-144 applied to every
                        # frame on its own, testing.yaml:357): an 80 ms frame at any rate becomes one 80 ms chunk at the
                        # model rate; the engine is shared with the scheduler thread, hence the lock
                        with self.lock:
                            pcm = self.asr.engine.resample(torch.as_tensor(pcm, dtype=torch.float32)[None].to(self.asr.engine.device), sr)[0].cpu().numpy()
                    sch.feed(slot, pcm)
                sch.close(slot)            # in order: after the last frame's tokens have been delivered
            except Exception as e:  # noqa: BLE001 -- reported to the client below
                err.append(e)
                sch.disconnect(slot)       # abort: free the slot right away

        threading.Thread(target=reader, daemon=True).start()
        out = sch.results(slot)
        while True:
            item = out.get()
            if item is None:
                break
            y_all, y_new = item
            text, reset = differ.step(y_all, y_new)
            if reset:
                sch.request_reset(slot)
            if text is not None:
                yield encode_transcript(text)
        if err:
            context.abort(grpc.StatusCode.INVALID_ARGUMENT, str(err[0]))


def add_servicer_to_server(servicer: ASRServicer, server):
    import grpc

    ident = lambda b: b  # noqa: E731 -- messages are (de)serialised by the servicer
    handlers = {
        "Transcribe": grpc.unary_unary_rpc_method_handler(servicer.Transcribe, request_deserializer=ident, response_serializer=ident),
        "TranscribeStream": grpc.stream_stream_rpc_method_handler(servicer.TranscribeStream, request_deserializer=ident,
                                                                  response_serializer=ident),
    }
    server.add_generic_rpc_handlers((grpc.method_handlers_generic_handler(SERVICE, handlers),))


def serve(asr, address="[::]:50051", n_streams=64, workers=None, **kw):
    """api-server.py:138-152.  One worker thread per connection is only a message pump here (the model runs on the scheduler
    thread), so the pool is sized to the number of stream slots instead of the reference's 4."""
    import grpc

    server = grpc.server(futures.ThreadPoolExecutor(max_workers=workers or (n_streams + 4)))
    servicer = ASRServicer(asr, n_streams=n_streams, **kw)
    add_servicer_to_server(servicer, server)
    port = server.add_insecure_port(address)
    server.start()
    return server, servicer, port
