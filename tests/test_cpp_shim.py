"""The header-style r8b:: front-end (include/r8b/CDSPResampler.h) compiled against libr8bgpu.so."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import oracle_util as ou

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "shim_demo")


def build_demo(pkg):
    src = os.path.join(ROOT, "tests", "cpp", "shim_demo.cpp")
    lib_dir = os.path.dirname(pkg.lib_path())
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(pkg.lib_path())):
        return EXE
    gxx = shutil.which("g++")
    if gxx is None:
        if os.path.exists(EXE):
            return EXE
        pytest.skip("no g++ and no prebuilt demo")
    subprocess.run([gxx, "-O1", "-std=c++11", "-I", os.path.join(ROOT, "include"), src, "-o", EXE,
                    "-L", lib_dir, "-lr8bgpu", "-Wl,-rpath," + lib_dir], check=True)
    return EXE


def test_shim_compiles_and_plans(pkg, ref):
    exe = build_demo(pkg)
    out = subprocess.run([exe, "--plan", "44100", "96000", "65536"], capture_output=True, text=True, check=True).stdout.split()
    r = ref.Resampler(44100.0, 96000.0, 65536, 2.0, 180.15)
    assert [int(out[0]), int(out[1]), int(out[2])] == [r.max_out_len, r.in_len_before_out_pos(0),
                                                       r.input_required_for_output(1000)]


def test_shim_failures_do_not_hang(pkg, ref):
    """A refused plan / a missing device must surface through isValid() and must never spin (the reference has no
    error channel; its loops in oneshot() / getInLenBeforeOutStart() assume process() always makes progress)."""
    exe = build_demo(pkg)
    out = subprocess.run([exe, "--failures"], capture_output=True, text=True, check=True, timeout=120).stdout.split("\n")
    valid, start, zeros, a, b = [int(t) for t in out[0].split()]
    assert (valid, start, zeros, a, b) == (0, 0, 64, 0, 0)
    before, after, start2 = [int(t) for t in out[1].split()]
    assert before == 1
    if pkg.device_count() < 1:
        assert (after, start2) == (0, 0)
    else:
        assert after == 1 and start2 == ref.Resampler(44100.0, 96000.0, 1024, 2.0, 180.15).in_len_before_out_start(0)


@pytest.mark.gpu
def test_shim_process_matches_reference(pkg, ref, tmp_path):
    exe = build_demo(pkg)
    n_ch, frames, block = 2, 20000, 4096
    x = ou.white_noise(n_ch, frames, 31)
    fin, fout = str(tmp_path / "in.f64"), str(tmp_path / "out.f64")
    x.tofile(fin)
    subprocess.run([exe, fin, fout, str(n_ch), str(frames), "44100", "96000", str(block)], check=True)
    raw = open(fout, "rb").read()
    pos = 0
    for c in range(n_ch):
        n = int(np.frombuffer(raw[pos:pos + 8], dtype=np.int64)[0])
        y = np.frombuffer(raw[pos + 8:pos + 8 + 8 * n], dtype=np.float64)
        pos += 8 + 8 * n
        r = ref.Resampler(44100.0, 96000.0, block, 2.0, 180.15)
        yr = np.concatenate([r.process(x[c, i:i + block]) for i in range(0, frames, block)])
        assert len(yr) == n
        m, rr = ou.parity_metrics(y, yr)
        assert m <= 32 * ou.EPS and rr <= 4 * ou.EPS


@pytest.mark.gpu
def test_pull_mode_equals_push_mode(pkg, ref, tmp_path):
    """README.md:132-146: real-time callers pull output; the helper must hand out exactly the push-mode stream."""
    exe = build_demo(pkg)
    n_ch, frames, block = 2, 30000, 2048
    x = ou.white_noise(n_ch, frames, 13)
    fin, fout = str(tmp_path / "in.f64"), str(tmp_path / "out.f64")
    x.tofile(fin)
    subprocess.run([exe, fin, fout, str(n_ch), str(frames), "44100", "96000", str(block), "--pull"], check=True)
    raw = open(fout, "rb").read()
    pos = 0
    for c in range(n_ch):
        n = int(np.frombuffer(raw[pos:pos + 8], dtype=np.int64)[0])
        y = np.frombuffer(raw[pos + 8:pos + 8 + 8 * n], dtype=np.float64)
        pos += 8 + 8 * n
        r = ref.Resampler(44100.0, 96000.0, block, 2.0, 180.15)
        yr = np.concatenate([r.process(x[c, i:i + block]) for i in range(0, frames, block)])
        assert len(yr) == n
        m, rr = ou.parity_metrics(y, yr)
        assert m <= 32 * ou.EPS and rr <= 4 * ou.EPS


@pytest.mark.gpu
def test_cpp_batch_object_drives_every_gpu(pkg, ref, tmp_path):
    """r8b::CDSPResamplerBatch(Device = -1) shards its channels over the visible GPUs behind the C-ABI; on a one-GPU box
    R8BGPU_FORCE_SHARDS exercises the same front."""
    exe = build_demo(pkg)
    n_ch, frames, block = 5, 20000, 4096
    x = ou.white_noise(n_ch, frames, 17)
    fin, fout = str(tmp_path / "in.f64"), str(tmp_path / "out.f64")
    x.tofile(fin)
    env = dict(os.environ)
    if pkg.device_count() < 2:
        env["R8BGPU_FORCE_SHARDS"] = "2"
    res = subprocess.run([exe, fin, fout, str(n_ch), str(frames), "44100", "96000", str(block), "--batch"], check=True,
                         capture_output=True, text=True, env=env)
    assert int(res.stdout) == max(2, min(pkg.device_count(), n_ch))
    raw = open(fout, "rb").read()
    pos = 0
    for c in range(n_ch):
        n = int(np.frombuffer(raw[pos:pos + 8], dtype=np.int64)[0])
        y = np.frombuffer(raw[pos + 8:pos + 8 + 8 * n], dtype=np.float64)
        pos += 8 + 8 * n
        r = ref.Resampler(44100.0, 96000.0, block, 2.0, 180.15)
        yr = np.concatenate([r.process(x[c, i:i + block]) for i in range(0, frames, block)])
        assert len(yr) == n
        m, rr = ou.parity_metrics(y, yr)
        assert m <= 32 * ou.EPS and rr <= 4 * ou.EPS


# ---- the reference's DLL interface (DLL/r8bsrc.h): libr8bsrc.so driven from plain C -------------------------------------
DLL_EXE = os.path.join(ROOT, "tests", "cpp", "dll_demo")


def build_dll_demo(pkg):
    lib_dir = os.path.dirname(pkg.lib_path())
    dll = os.path.join(lib_dir, "libr8bsrc.so")
    if not os.path.exists(dll):
        pytest.skip("libr8bsrc.so not built (no g++ at build time)")
    src = os.path.join(ROOT, "tests", "cpp", "dll_demo.c")
    if os.path.exists(DLL_EXE) and os.path.getmtime(DLL_EXE) > max(os.path.getmtime(src), os.path.getmtime(dll)):
        return DLL_EXE
    gcc = shutil.which("gcc")
    if gcc is None:
        if os.path.exists(DLL_EXE):
            return DLL_EXE
        pytest.skip("no gcc and no prebuilt demo")
    subprocess.run([gcc, "-O1", "-std=c99", "-I", os.path.join(ROOT, "include"), src, "-o", DLL_EXE,
                    "-L", lib_dir, "-lr8bsrc", "-lr8bgpu", "-Wl,-rpath," + lib_dir], check=True)
    return DLL_EXE


def test_dll_interface_exports_and_inlen(pkg, ref):
    exe = build_dll_demo(pkg)
    import ctypes
    L = ctypes.CDLL(os.path.join(os.path.dirname(pkg.lib_path()), "libr8bsrc.so"))
    for name in ("r8b_create", "r8b_delete", "r8b_inlen", "r8b_clear", "r8b_process"):  # DLL/r8bsrc.h:71-134
        assert hasattr(L, name)
    for res, atten in ((0, 136.45), (1, 109.56), (2, 180.15)):
        out = subprocess.run([exe, "--inlen", "44100", "96000", "4096", str(res), "1000"], capture_output=True, text=True, check=True)
        assert int(out.stdout) == ref.Resampler(44100.0, 96000.0, 4096, 2.0, atten).input_required_for_output(1000)


@pytest.mark.gpu
def test_dll_interface_process_matches_reference(pkg, ref, tmp_path):
    exe = build_dll_demo(pkg)
    frames, block = 30000, 4096
    x = ou.white_noise(1, frames, 9)[0]
    fin, fout = str(tmp_path / "in.f64"), str(tmp_path / "out.f64")
    x.tofile(fin)
    for res, atten in ((0, 136.45), (2, 180.15)):
        out = subprocess.run([exe, fin, fout, str(frames), "48000", "44100", str(block), str(res)], capture_output=True, text=True, check=True)
        y = np.fromfile(fout, dtype=np.float64)
        # the demo clears after the first block and feeds it again: the stream restarts
        r = ref.Resampler(48000.0, 44100.0, block, 2.0, atten)
        parts = [r.process(x[:block])]          # written before the clear
        r.clear()
        r.process(x[:block])                    # re-fed after the clear (not written)
        parts += [r.process(x[i:i + block]) for i in range(block, frames, block)]
        yr = np.concatenate(parts)
        assert int(out.stdout) == len(yr) == len(y)
        m, rr = ou.parity_metrics(y, yr)
        assert m <= 32 * ou.EPS and rr <= 4 * ou.EPS
