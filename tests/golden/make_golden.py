#!/usr/bin/env python3
"""Generate the committed golden fixtures from the REFERENCE ITSELF (dev container only).

Needs /root/reference (for the WAV pair) and oracle/_ref/libr8bref_e{0,1}.so (the unmodified
reference headers compiled by oracle/Makefile).  Outputs, all small:

  ref_vectors.npz   seeded fp64 inputs, per-call output counts and outputs of r8b::CDSPResampler
                    (CDSPResampler24 unless noted) for the BASELINE chains + planner branches.
                    Long outputs are stored strided together with sum / sum-of-squares.
  drums_excerpt.npz first 0.5 s of bench/DrumsSrc.wav (24-bit, 2 ch) and the matching prefix of the
                    author's own conversion bench/DrumsDst96.wav -- the reference's only KAT.
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_util as ou  # noqa: E402

REF = os.environ.get("R8B_REFERENCE", "/root/reference")

CASES = [
    # name, src, dst, tb, atten, extfft, block lens
    ("cfg1_44100_96000", 44100.0, 96000.0, 2.0, 180.15, 0, [2048, 2048, 2048, 1000, 1, 0, 17, 2048]),
    ("cfg3_48000_44100", 48000.0, 44100.0, 2.0, 180.15, 0, [2048] * 4),
    ("cfg5_48000_47999", 48000.0, 47999.0, 2.0, 180.15, 0, [2048] * 4 + [333, 2048]),
    ("cfg4_dsd_extfft", 44100.0, 2822400.0, 2.0, 180.15, 1, [2048] * 3),
    ("hbdown_192000_44100", 192000.0, 44100.0, 2.0, 180.15, 0, [4096] * 4),
    ("half_96000_48000", 96000.0, 48000.0, 2.0, 180.15, 0, [4096] * 3),
    ("third_48000_16000", 48000.0, 16000.0, 2.0, 180.15, 0, [4096] * 3),
    ("interm_44100_192000", 44100.0, 192000.0, 2.0, 180.15, 0, [2048] * 4),
    ("r16_44100_48000", 44100.0, 48000.0, 3.0, 136.45, 0, [2048] * 4),
    ("up3_32000_48000", 32000.0, 48000.0, 2.0, 180.15, 0, [2048] * 4),
    ("up6_8000_48000", 8000.0, 48000.0, 2.0, 180.15, 0, [2048] * 3),
]
MAX_STORE = 8192


def read_wav24(path):
    b = open(path, "rb").read()
    assert b[:4] == b"RIFF" and b[8:12] == b"WAVE"
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(b):
        cid, sz = b[pos:pos + 4], struct.unpack("<I", b[pos + 4:pos + 8])[0]
        if cid == b"fmt ":
            fmt = struct.unpack("<HHIIHH", b[pos + 8:pos + 24])
        elif cid == b"data":
            data = b[pos + 8:pos + 8 + sz]
        pos += 8 + sz + (sz & 1)
    tag, ch, rate, _, _, bits = fmt
    assert bits == 24
    raw = np.frombuffer(data, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
    v = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.int32)
    return rate, v.reshape(-1, ch)


def main():
    out = {}
    names = []
    for name, src, dst, tb, att, ext, lens in CASES:
        ref = ou.RefOracle("e1" if ext else "e0")
        x = ou.white_noise(1, int(sum(lens)), seed=abs(hash(name)) % 1000 + 1)[0] if False else \
            np.random.default_rng([len(name), int(src), int(dst)]).uniform(-1, 1, int(sum(lens)))
        r = ref.Resampler(src, dst, max(lens), tb, att)
        pos, ys, counts = 0, [], []
        for l in lens:
            y = r.process(x[pos:pos + l])
            pos += l
            ys.append(y)
            counts.append(len(y))
        y = np.concatenate(ys)
        stride = max(1, -(-len(y) // MAX_STORE))
        names.append(name)
        out[name + "/params"] = np.array([src, dst, tb, att, ext, stride, r.max_out_len,
                                          r.in_len_before_out_pos(0), r.in_len_before_out_pos(1000)], dtype=np.float64)
        out[name + "/lens"] = np.array(lens, dtype=np.int64)
        out[name + "/counts"] = np.array(counts, dtype=np.int64)
        out[name + "/x"] = x
        out[name + "/y_sub"] = y[::stride].copy()
        out[name + "/y_stats"] = np.array([len(y), y.sum(), (y * y).sum(), np.abs(y).max()])
        print(name, "in", len(x), "out", len(y), "stride", stride, counts)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)

    rate_s, s = read_wav24(os.path.join(REF, "bench", "DrumsSrc.wav"))
    rate_d, d = read_wav24(os.path.join(REF, "bench", "DrumsDst96.wav"))
    assert (rate_s, rate_d) == (44100, 96000), (rate_s, rate_d)
    n_src = 22050
    n_dst = int((n_src - 400) * 96000 // 44100)
    np.savez_compressed(os.path.join(HERE, "drums_excerpt.npz"), src=s[:n_src].copy(), dst=d[:n_dst].copy(),
                        full_frames=np.array([len(s), len(d)]))
    print("drums: src", s.shape, "dst", d.shape, "excerpt", n_src, n_dst)


if __name__ == "__main__":
    main()
