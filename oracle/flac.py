"""TEST INFRASTRUCTURE ONLY -- a minimal FLAC decoder (no FLAC library is installed and there
is no network).  Used by ``oracle/make_golden.py`` to read the reference's demo utterance
(``demo/3729-6852-0035.flac``, api-client.py:13) for the BASELINE config-1 fixture.  Written from
the public FLAC format specification; supports what that file needs and a bit more: fixed block
size streams, 8-24 bit mono / independent channels, CONSTANT / VERBATIM / FIXED / LPC subframes,
partitioned Rice residuals (both parameter widths, escape codes).  The decode is verified against
the STREAMINFO MD5 of the PCM (93b7bac14eaaf8c6b0ad3b3be8c63c24)."""
import hashlib
import struct

import numpy as np


class _Bits:
    def __init__(self, data, pos):
        self.d, self.p, self.acc, self.n = data, pos, 0, 0

    def read(self, k):
        while self.n < k:
            self.acc = (self.acc << 8) | self.d[self.p]
            self.p += 1
            self.n += 8
        self.n -= k
        v = (self.acc >> self.n) & ((1 << k) - 1)
        self.acc &= (1 << self.n) - 1
        return v

    def read_signed(self, k):
        v = self.read(k)
        return v - (1 << k) if v >> (k - 1) else v

    def unary(self):
        z = 0
        while self.read(1) == 0:
            z += 1
        return z

    def align(self):
        self.n -= self.n % 8
        self.acc &= (1 << self.n) - 1

    def byte_pos(self):
        return self.p - self.n // 8


def _residual(br, blocksize, order, out):
    method = br.read(2)
    pbits = 4 if method == 0 else 5
    esc = (1 << pbits) - 1
    porder = br.read(4)
    nparts = 1 << porder
    i = order
    for part in range(nparts):
        cnt = (blocksize >> porder) - (order if part == 0 else 0)
        k = br.read(pbits)
        if k == esc:
            nb = br.read(5)
            for _ in range(cnt):
                out[i] = br.read_signed(nb) if nb else 0
                i += 1
        else:
            for _ in range(cnt):
                q = br.unary()
                r = br.read(k) if k else 0
                u = (q << k) | r
                out[i] = (u >> 1) ^ -(u & 1)
                i += 1


_FIXED = {0: (), 1: (1,), 2: (2, -1), 3: (3, -3, 1), 4: (4, -6, 4, -1)}


def _subframe(br, blocksize, bps):
    assert br.read(1) == 0
    typ = br.read(6)
    wasted = 0
    if br.read(1):
        wasted = br.unary() + 1
        bps -= wasted
    s = [0] * blocksize
    if typ == 0:  # CONSTANT
        s = [br.read_signed(bps)] * blocksize
    elif typ == 1:  # VERBATIM
        s = [br.read_signed(bps) for _ in range(blocksize)]
    elif 8 <= typ <= 12:  # FIXED
        order = typ - 8
        for i in range(order):
            s[i] = br.read_signed(bps)
        _residual(br, blocksize, order, s)
        co = _FIXED[order]
        for i in range(order, blocksize):
            s[i] += sum(c * s[i - 1 - j] for j, c in enumerate(co))
    elif typ >= 32:  # LPC
        order = (typ & 31) + 1
        for i in range(order):
            s[i] = br.read_signed(bps)
        prec = br.read(4) + 1
        shift = br.read_signed(5)
        co = [br.read_signed(prec) for _ in range(order)]
        _residual(br, blocksize, order, s)
        for i in range(order, blocksize):
            acc = 0
            for j in range(order):
                acc += co[j] * s[i - 1 - j]
            s[i] += acc >> shift
    else:
        raise ValueError(f"reserved subframe type {typ}")
    if wasted:
        s = [v << wasted for v in s]
    return s


_BS = {1: 192, 2: 576, 3: 1152, 4: 2304, 5: 4608}


def decode(path):
    """Returns (pcm int32 [channels, n], sample_rate, bits_per_sample); checks the STREAMINFO MD5."""
    d = open(path, "rb").read()
    assert d[:4] == b"fLaC"
    pos, info = 4, None
    while True:
        hdr = d[pos]
        ln = int.from_bytes(d[pos + 1:pos + 4], "big")
        if hdr & 0x7F == 0:
            info = d[pos + 4:pos + 4 + ln]
        pos += 4 + ln
        if hdr & 0x80:
            break
    x = int.from_bytes(info[10:18], "big")
    sr, ch, bps, total = x >> 44, ((x >> 41) & 7) + 1, ((x >> 36) & 31) + 1, x & ((1 << 36) - 1)
    md5 = info[18:34]
    chans = [[] for _ in range(ch)]
    while pos < len(d) and len(chans[0]) < total:
        br = _Bits(d, pos)
        assert br.read(14) == 0x3FFE, "lost frame sync"
        br.read(1)
        br.read(1)  # blocking strategy
        bsc, src = br.read(4), br.read(4)
        chan_assign, ssc = br.read(4), br.read(3)
        br.read(1)
        first = br.read(8)  # UTF-8 coded frame/sample number
        extra = 0
        while first & 0x80 and (first << extra) & 0x40:
            extra += 1
        if first & 0x80:
            n_more = 1
            t = first << 1
            while t & 0x80:
                n_more += 1
                t <<= 1
            for _ in range(n_more - 1):
                br.read(8)
        if bsc == 6:
            blocksize = br.read(8) + 1
        elif bsc == 7:
            blocksize = br.read(16) + 1
        elif bsc >= 8:
            blocksize = 256 << (bsc - 8)
        else:
            blocksize = _BS[bsc]
        if src == 12:
            br.read(8)
        elif src in (13, 14):
            br.read(16)
        br.read(8)  # CRC-8
        fbps = {0: bps, 1: 8, 2: 12, 4: 16, 5: 20, 6: 24}[ssc]
        if chan_assign < 8:
            subs = [_subframe(br, blocksize, fbps) for _ in range(chan_assign + 1)]
        else:  # stereo decorrelation
            if chan_assign == 8:
                l = _subframe(br, blocksize, fbps); sd = _subframe(br, blocksize, fbps + 1)
                subs = [l, [a - b for a, b in zip(l, sd)]]
            elif chan_assign == 9:
                sd = _subframe(br, blocksize, fbps + 1); r = _subframe(br, blocksize, fbps)
                subs = [[a + b for a, b in zip(sd, r)], r]
            else:
                m = _subframe(br, blocksize, fbps); sd = _subframe(br, blocksize, fbps + 1)
                subs = [[((2 * a + (b & 1)) + b) >> 1 for a, b in zip(m, sd)], [((2 * a + (b & 1)) - b) >> 1 for a, b in zip(m, sd)]]
        br.align()
        br.read(16)  # CRC-16
        pos = br.byte_pos()
        for c in range(ch):
            chans[c].extend(subs[c])
    pcm = np.asarray(chans, dtype=np.int32)[:, :total]
    nbytes = (bps + 7) // 8
    inter = pcm.T.reshape(-1)
    raw = b"".join(int(v).to_bytes(nbytes, "little", signed=True) for v in inter) if nbytes != 2 else inter.astype("<i2").tobytes()
    if hashlib.md5(raw).digest() != md5:
        raise ValueError("FLAC decode does not match the STREAMINFO MD5")
    return pcm, sr, bps
