"""CPU tests of the serving wire path (libreasr_b200/serve.py): message bytes against google.protobuf built from the
reference's .proto definition, the WebSocket bridge frame, the transcript post-processing of api-server.py:117-135
against a literal transcription of that loop, the connection scheduler against a fake streaming session, and the gRPC
service end to end over loopback with a fake engine (the real engine is exercised by the GPU test)."""
import itertools as it
import struct
import threading
import time
from concurrent import futures

import numpy as np
import pytest
import torch

from libreasr_b200 import serve as S


# ---- protobuf messages exactly as interfaces/libreasr.proto:10-17 declares them ----
def _proto_classes():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

    fd = descriptor_pb2.FileDescriptorProto(name="libreasr_test.proto", package="ASR", syntax="proto3")
    a = fd.message_type.add(name="Audio")
    a.field.add(name="data", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_BYTES,
                label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    a.field.add(name="sr", number=3, type=descriptor_pb2.FieldDescriptorProto.TYPE_INT32,
                label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    t = fd.message_type.add(name="Transcript")
    t.field.add(name="data", number=1, type=descriptor_pb2.FieldDescriptorProto.TYPE_STRING,
                label=descriptor_pb2.FieldDescriptorProto.LABEL_OPTIONAL)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = getattr(message_factory, "GetMessageClass", None)
    mk = (lambda d: get(d)) if get else (lambda d: message_factory.MessageFactory(pool).GetPrototype(d))
    return mk(pool.FindMessageTypeByName("ASR.Audio")), mk(pool.FindMessageTypeByName("ASR.Transcript"))


@pytest.mark.parametrize("n,sr", [(0, 0), (5, 16000), (1280, 16000), (40000, 44100), (3, 8000), (2, -1), (1, -2**31), (1, 2**31 - 1)])
def test_audio_message_bytes_match_protobuf(n, sr):
    Audio, _ = _proto_classes()
    data = np.random.default_rng(n).standard_normal(n).astype("<f4").tobytes()
    ref = Audio(data=data, sr=sr).SerializeToString()
    assert S.encode_audio(data, sr) == ref
    assert S.decode_audio(ref) == (data, sr)
    m = Audio()
    m.ParseFromString(S.encode_audio(data, sr))
    assert (m.data, m.sr) == (data, sr)


@pytest.mark.parametrize("text", ["", "hello", "grüße ▁wörld", "x" * 300])
def test_transcript_message_bytes_match_protobuf(text):
    _, Transcript = _proto_classes()
    ref = Transcript(data=text).SerializeToString()
    assert S.encode_transcript(text) == ref
    assert S.decode_transcript(ref) == text


def test_ws_bridge_frame_round_trip():
    pcm = np.linspace(-1, 1, 1280, dtype=np.float32)
    frame = S.build_ws_frame("en", 16000, pcm)
    assert frame[:4] == b"en  " and struct.unpack("f", frame[4:8])[0] == 16000.0      # api-bridge.py:97-102
    lang, sr, data = S.parse_ws_frame(frame)
    assert (lang, sr) == ("en", 16000)
    np.testing.assert_array_equal(S.tensorize(data)[0], pcm)
    with pytest.raises(ValueError):
        S.parse_ws_frame(b"en")


def _reference_postprocess(outputs, denum, downsample, n_buffer):
    """api-server.py:117-135 transcribed literally (outputs: iterable of (y, y_one_text)); returns (sent, reset steps)."""
    sent, resets = [], []
    last, last_diff, steps = "", "", 0
    for i, (y, y_one) in enumerate(outputs):
        steps += 1
        if y_one != "":
            now = denum(y)
            diff = "".join(b for a, b in it.zip_longest(last, now) if a != b)
            last = now
            if diff == last_diff:
                continue
            last_diff = diff
            sent.append(diff)
        elif int(10.0 * downsample * n_buffer * steps) >= 4000:
            resets.append(i)
            steps = 0
    return sent, resets


def test_transcript_differ_equals_reference_loop():
    denum = lambda ids: "".join(chr(97 + t % 26) for t in ids)   # noqa: E731
    rng = np.random.default_rng(3)
    y, outs = [], []
    for _ in range(120):
        new = list(rng.integers(0, 26, size=rng.choice([0, 0, 0, 1, 2]))) if rng.random() < 0.6 else []
        if rng.random() < 0.1 and y:
            new = [y[-1]] if rng.random() < 0.5 else new   # repeated characters -> repeated diffs
        y = y + new
        outs.append((list(y), list(new)))
    want_sent, want_resets = _reference_postprocess([(yy, denum(nn)) for yy, nn in outs], denum, 8, 2)
    d = S.TranscriptDiffer(denum, 8, 2)
    sent, resets = [], []
    for i, (yy, nn) in enumerate(outs):
        text, reset = d.step(yy, nn)
        if text is not None:
            sent.append(text)
        if reset:
            resets.append(i)
    assert (sent, resets) == (want_sent, want_resets)
    assert len(resets) >= 1 and len(sent) >= 10


class FakeSession:
    """Stands in for StreamBatch: stream b's model step (every 2nd chunk after 2 warm-up chunks, like the reference window +
    Buffer) emits one token = number of chunks seen, offset by 100 * b, so routing mistakes show up in the ids."""

    def __init__(self, B, chunk):
        self.B, self.chunk = B, chunk
        self.seen = [0] * B
        self.rows = [0] * B
        self.resets = []
        self.ran = [False] * B

    def reset(self, slot):
        self.seen[slot] = self.rows[slot] = 0
        self.resets.append(slot)

    def push(self, chunks, active=None):
        assert tuple(chunks.shape) == (self.B, self.chunk)
        new, self.ran, any_ran = [[] for _ in range(self.B)], [False] * self.B, False
        for b in range(self.B):
            if active is not None and not active[b]:
                continue
            self.seen[b] += 1
            if self.seen[b] < 3:
                continue
            self.rows[b] += 1
            if self.rows[b] == 2:
                self.rows[b] = 0
                self.ran[b] = any_ran = True
                if float(chunks[b, 0]) >= 0:       # "silent" frames (negative marker) emit nothing
                    new[b] = [100 * b + self.seen[b]]
        return new if any_ran else None


def test_scheduler_routes_streams_independently():
    ses = FakeSession(4, 8)
    sch = S.StreamScheduler(ses)
    a = sch.connect()
    for _ in range(6):
        sch.feed(a, np.zeros(8, np.float32))
    while sch.tick():
        pass
    b = sch.connect()                                  # second connection starts later, own phase
    for _ in range(4):
        sch.feed(b, np.zeros(8, np.float32))
        sch.feed(a, np.zeros(8, np.float32))
    while sch.tick():
        pass
    ra = [sch.results(a).get_nowait() for _ in range(sch.results(a).qsize())]
    rb = [sch.results(b).get_nowait() for _ in range(sch.results(b).qsize())]
    assert [n for _, n in ra] == [[100 * a + 4], [100 * a + 6], [100 * a + 8], [100 * a + 10]]
    assert [n for _, n in rb] == [[100 * b + 4]]
    assert ra[-1][0] == [100 * a + 4, 100 * a + 6, 100 * a + 8, 100 * a + 10]
    assert ses.resets == [a, b]                        # a slot is reset when a connection takes it
    sch.disconnect(a)
    assert sch.results(a).get_nowait() is None
    c = sch.connect()                                  # the freed slot is re-used and reset again
    sch.feed(c, np.zeros(8, np.float32))
    sch.tick()
    assert c == a and ses.resets == [a, b, a]
    with pytest.raises(ValueError):
        sch.feed(c, np.zeros(5, np.float32))
    for _ in range(2):
        sch.connect()
    with pytest.raises(RuntimeError):
        sch.connect()


class FakeASR:
    class _Cfg:
        sample_rate = 16000

    class _Eng:
        pass

    def __init__(self):
        self.engine = FakeASR._Eng()
        self.engine.cfg = FakeASR._Cfg()

    def transcribe(self, audio, sr=16000):
        return [int(audio.numel()), int(sr)]


def test_grpc_service_over_loopback_with_fake_engine():
    grpc = pytest.importorskip("grpc")
    denum = lambda ids: "".join(chr(97 + t % 26) for t in ids)   # noqa: E731
    asr = FakeASR()
    sch = S.StreamScheduler(FakeSession(3, 16))
    servicer = S.ASRServicer(asr, scheduler=sch, denumericalize=denum)
    server = grpc.server(futures.ThreadPoolExecutor(max_workers=8))
    S.add_servicer_to_server(servicer, server)
    port = server.add_insecure_port("127.0.0.1:0")
    server.start()
    try:
        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        ident = lambda b: b  # noqa: E731
        unary = ch.unary_unary(f"/{S.SERVICE}/Transcribe", request_serializer=ident, response_deserializer=ident)
        pcm = np.zeros(320, np.float32)
        assert S.decode_transcript(unary(S.encode_audio(pcm.tobytes(), 8000), timeout=10)) == denum([320, 8000])
        stream = ch.stream_stream(f"/{S.SERVICE}/TranscribeStream", request_serializer=ident, response_deserializer=ident)

        def client(n_frames, out):
            def gen():
                for _ in range(n_frames):
                    yield S.encode_audio(np.zeros(16, np.float32).tobytes(), 16000)
                    time.sleep(0.002)
            out.extend(S.decode_transcript(r) for r in stream(gen(), timeout=20))

        outs = [[], []]
        ths = [threading.Thread(target=client, args=(10, outs[0])), threading.Thread(target=client, args=(6, outs[1]))]
        for t in ths:
            t.start()
        for t in ths:
            t.join(timeout=30)
        # each connection ran steps at its chunks 4, 6, 8, 10 (resp. 4, 6): one new character per step, sent as the diff
        assert [len(o) for o in outs] == [4, 2]
        assert all(len(x) == 1 for o in outs for x in o)
        bad = stream(iter([S.encode_audio(np.zeros(5, np.float32).tobytes(), 16000)]), timeout=10)
        with pytest.raises(grpc.RpcError):
            list(bad)
    finally:
        server.stop(0)
        servicer.close()
