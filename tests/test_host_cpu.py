"""CPU-only tests: the C-ABI library loads and exports everything include/gf3_b200.h declares,
the product fails loudly without a CUDA device (no CPU fallback), host-side logic of the drop-in
module, and the sharded-sweep reduction over gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import PKG_DIR, ROOT, load_golden
from oracle import gf3_oracle as orc


def _header_functions():
    text = open(os.path.join(ROOT, "include", "gf3_b200.h")).read()
    return sorted(set(re.findall(r"GF3_API[^;(]*?\b(gf3_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from gf3b200 import _lib
    names = _header_functions()
    assert len(names) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libgf3b200.so does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, "ctypes SIGNATURES out of sync with the header"
    assert _lib.load().gf3_abi_version() == 1
    # the library exports nothing else
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T " in l)
    assert exported == names


def test_params_default_and_struct_layout():
    from gf3b200 import _lib
    lib = _lib.load()
    p = _lib.Gf3Params()
    _lib.check(lib.gf3_params_default(ctypes.byref(p), 4096, 224, 100, 1500, 20, 180))
    assert (p.N, p.cp, p.lo, p.hi, p.n_pilots, p.packet_len) == (4096, 224, 100, 1500, 20, 180)
    assert (p.fit_lo, p.fit_hi, p.chirp_len) == (500, 1000, 21600)                # OFDM.py:462,64
    assert (p.fs, p.f0, p.f1) == (48000.0, 0.0, 8000.0) and abs(p.thresh - 0.4) < 1e-7
    assert p.tx_gain == 2.0 and abs(p.chirp_gain - 0.2) < 1e-7
    assert ctypes.sizeof(_lib.Gf3Params) == 15 * 4


def test_no_cpu_fallback_without_device():
    """On a box without a GPU every compute path must raise, never silently compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    import gf3b200
    from gf3b200 import _lib
    lib = _lib.load()
    assert lib.gf3_device_count() == 0
    p = _lib.Gf3Params()
    _lib.check(lib.gf3_params_default(ctypes.byref(p), 1024, 32, 1, 512, 20, 180))
    plan = ctypes.c_void_p()
    rc = lib.gf3_plan_create(ctypes.byref(p), ctypes.byref(plan))
    assert rc == _lib.GF3_ERR_NODEVICE and b"no CPU path" in lib.gf3_last_error()
    with pytest.raises(gf3b200.Gf3Error):
        gf3b200.Phy(N=1024, cp=32, lo=1, hi=512)
    import OFDM
    rx = OFDM.receiver("A2", "XOR")
    with pytest.raises(gf3b200.Gf3Error):
        rx.receive(np.zeros(100000))
    with pytest.raises(gf3b200.Gf3Error):
        OFDM.transmitter("A2", "XOR").transmit(np.zeros(100, dtype=np.int64))


def test_invalid_params_rejected():
    from gf3b200 import _lib
    lib = _lib.load()
    for bad in [(1000, 32, 1, 400, 20, 180), (1024, -1, 1, 512, 20, 180), (1024, 32, 0, 512, 20, 180),
                (1024, 32, 1, 513, 20, 180), (8192, 32, 1, 512, 20, 180), (1024, 32, 1, 512, 20, 0)]:
        p = _lib.Gf3Params()
        lib.gf3_params_default(ctypes.byref(p), *bad)
        plan = ctypes.c_void_p()
        rc = lib.gf3_plan_create(ctypes.byref(p), ctypes.byref(plan))
        assert rc in (_lib.GF3_ERR_INVALID, _lib.GF3_ERR_NODEVICE) and rc != 0
        if rc == _lib.GF3_ERR_INVALID:
            assert len(lib.gf3_last_error()) > 0


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under gf3-audio-modem_b200/ may reference it."""
    for root, _, files in os.walk(PKG_DIR):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "oracle" not in text.replace("the numpy oracle", ""), "%s mentions the oracle" % f


def test_dropin_parameter_object_matches_reference_contract(known_sequence):
    import OFDM
    for mode, (cp, (lo, hi)) in orc.MODES.items():
        c = OFDM.CamG(mode, encoding="XOR")
        p = orc.Params.from_mode(mode, known_sequence=known_sequence)
        assert (c.ofdm_symbol_size, c.K, c.cp_length, c.lowest_bin, c.highest_bin) == (4096, 2047, cp, lo, hi)
        assert c.chirp_length == p.chirp_length and c.data_bits_per_symbol == p.data_bits_per_symbol
        assert c.bits_per_symbol == 4094 and c.no_pilots == 20 and c.packet_length == 180
        assert np.array_equal(c.data_carriers, p.data_carriers) and np.array_equal(c.unused_carriers, p.unused_carriers)
        assert np.array_equal(c.known_sequence, known_sequence)
    assert "Number of actual Sub Carriers:      2047" in repr(OFDM.CamG("A2"))
    with pytest.raises(KeyError):
        OFDM.CamG("Z9")
    c = OFDM.CamG("A1", ofdm_symbol_size=1024, cp_length=32, lowest_bin=1, highest_bin=512)
    assert (c.K, c.chirp_length, c.data_carriers_per_symbol) == (511, 5280, 511)


def test_dropin_encode_decode_match_oracle(known_sequence):
    import OFDM
    tx = OFDM.transmitter("A2", "XOR", no_pilots=4, packet_length=16)
    p = orc.Params.from_mode("A2", known_sequence=known_sequence, n_pilots=4, packet_len=16)
    bits = np.random.default_rng(1).integers(0, 2, 50001)
    np.random.seed(3)
    a = tx.encode(bits)
    np.random.seed(3)
    b = orc.encode(p, bits)
    assert np.array_equal(a, b) and len(a) % (2800 * 16) == 0
    rx = OFDM.receiver("A2", "XOR", no_pilots=4, packet_length=16)
    assert np.array_equal(rx.decode(a)[: len(bits)], bits)
    assert np.array_equal(rx.PS(tx.SP(a)), a) and tx.SP(a).shape[1:] == (1400, 2)
    assert np.allclose(tx.map(tx.SP(a))[:3], orc.qpsk_map(a.reshape(-1, 1400, 2))[:3])
    np.random.seed(4)
    f1 = tx.random_qpsk()
    np.random.seed(4)
    f2 = orc.random_qpsk(p)
    assert np.array_equal(f1, f2)
    sym = tx.build_OFDM_symbol(tx.map(tx.SP(a))[:2])
    assert sym.shape == (2, 4096) and np.allclose(sym[:, 1:2048], np.conj(sym[:, :2048:-1]))
    assert tx.add_cp(np.arange(8192.0).reshape(2, 4096)).shape == (2, 4320)
    with pytest.raises(NotImplementedError):
        OFDM.transmitter("A2", "LDPC").encode(bits)


def test_dropin_file_framing(tmp_path, monkeypatch):
    """load_file / save_file (OFDM.py:756-794) incl. the case-insensitive directory lookup."""
    import OFDM
    g = load_golden("kat1_gr5ch1.npz")
    (tmp_path / "input_Files").mkdir()
    (tmp_path / "output_files").mkdir()
    g["bmp"].tofile(tmp_path / "input_Files" / "gr5ch1.bmp")
    monkeypatch.chdir(tmp_path)
    bits = OFDM.load_file("gr5ch1.bmp")
    assert np.array_equal(bits, orc.load_file_bits("gr5ch1.bmp", g["bmp"])) and len(bits) == 1049168
    name, data = OFDM.save_file(np.concatenate([bits, np.ones(77, dtype=np.uint8)]))
    assert name == "gr5ch1.bmp" and np.array_equal(data, g["bmp"])
    assert np.array_equal(np.fromfile(tmp_path / "output_files" / "gr5ch1_received.bmp", dtype=np.uint8), g["bmp"])
    ref_bits = np.unpackbits(g["bits_packed"])[:1512000]
    name2, data2 = OFDM.save_file(ref_bits)                  # the reference's own decoded bits (2.3 % BER)
    assert name2 == str(g["file_name"]) and np.array_equal(data2, g["file_payload"])
    # the packed-byte forms carry the same stream as bytes
    fb = OFDM.load_file_bytes("gr5ch1.bmp")
    assert fb.dtype == np.uint8 and np.array_equal(np.unpackbits(fb), bits)
    name3, data3 = OFDM.save_file_bytes(np.concatenate([fb, np.full(9, 255, dtype=np.uint8)]))
    assert name3 == "gr5ch1.bmp" and np.array_equal(data3, g["bmp"])


def test_get_symbols_slicing_matches_oracle(known_sequence):
    import OFDM
    g = load_golden("stage_w1024.npz")
    N, cp, lo, hi, P, L, npk = (int(v) for v in g["cfg"])
    rx = OFDM.receiver("A1", "XOR", no_pilots=P, packet_length=L, ofdm_symbol_size=N, cp_length=cp, lowest_bin=lo, highest_bin=hi)
    p = orc.Params(N=N, cp=cp, lo=lo, hi=hi, n_pilots=P, packet_len=L, known_sequence=known_sequence)
    r = g["r_i16"].astype(np.float64)
    zeros = np.zeros(len(r) + p.chirp_length - 3, dtype=bool)
    zeros[g["peaks"]] = True
    a = rx.get_symbols(r, zeros)
    b, starts = orc.get_symbols(p, r, zeros)
    assert np.array_equal(a, b) and rx.no_packets == npk == len(starts)
    assert rx.remove_cp(a).shape == (npk, 2 * P + L, N)
    d, s, e = rx.get_data(np.fft.fft(rx.remove_cp(a)))
    assert d.shape == (npk, L, N // 2 - 1) and s.shape == e.shape == (npk, P, N // 2 - 1)


def test_shard_streams_partition():
    from gf3b200.sweep import shard_streams
    for n, w in [(10, 1), (10, 3), (4096, 8), (5, 8)]:
        parts = [shard_streams(n, r, w) for r in range(w)]
        assert sorted(np.concatenate(parts).tolist()) == list(range(n))
        assert all(np.all(p % w == r) for r, p in enumerate(parts))


_GLOO_WORKER = r'''
import os, sys
sys.path.insert(0, {pkg!r})
import numpy as np, torch, torch.distributed as dist
from gf3b200.sweep import sweep
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=rank, world_size=world)
def count(ids, snr):      # deterministic per-stream fake counters
    ids = np.asarray(ids)
    return int(np.sum((ids * 7 + int(snr)) % 13)), int(len(ids) * 1000), int(np.sum(ids % 5 == 0))
res = sweep(count, 37, [0, 4, 8], rank, world, dist, chunk=4)
ref = sweep(count, 37, [0, 4, 8], 0, 1, None, chunk=9)
assert np.array_equal(res, ref), (res, ref)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_sharded_sweep_gloo_world2(tmp_path):
    """N > 1 path on CPU: two gloo ranks, stream_id % 2 sharding, all-reduced counters equal the
    unsharded ones."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "w.py"
    script.write_text(_GLOO_WORKER.format(pkg=PKG_DIR, port=port))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for pr in procs:
        out, _ = pr.communicate(timeout=120)
        assert pr.returncode == 0, out
        assert "ok" in out


def test_numa_placement_helper_is_harmless_without_a_gpu():
    """bind_to_gpu_numa (bench.py calls it before pinning host buffers) reports what it found and never raises:
    no NVML / no device here, so it must come back with an error note and change nothing."""
    import os
    from gf3b200.host import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    info = bind_to_gpu_numa(0)
    assert isinstance(info, dict) and info["gpu"] == 0
    assert os.sched_getaffinity(0) == before or info.get("cpus_bound")


def test_old_api_parameter_objects_without_a_device():
    """The old-API constructors only build parameter objects (no device needed until a stage method runs)."""
    import OFDM
    wc = OFDM.CamG(1024, 32, "QPSK")
    assert (wc.K, wc.cp_length, wc.bits_per_symbol, len(wc.all_carriers)) == (1024, 32, 1022, 1024)
    x = np.arange(2 * 1056.0).reshape(2, 1056)
    assert wc.remove_cp(x).shape == (2, 1024) and np.array_equal(wc.add_cp(wc.remove_cp(x))[:, 32:], x[:, 32:])
    bits = np.array([0, 0, 1, 0, 1, 1, 0, 1])
    assert np.allclose(wc.map(wc.SP(bits)) * np.sqrt(2), [1 + 1j, 1 - 1j, -1 - 1j, -1 + 1j])
    rx = OFDM.receiver(ofdm_symbol_size=4096, cp_length=0, modulation="QPSK", fs=48000, end_sync=False)
    assert (rx.K, rx.cp_length, rx.data_carriers_per_symbol, rx.old_api, rx.end_sync) == (2047, 0, 2047, True, False)
    tx = OFDM.transmitter(1024, 128, "QPSK")
    assert (tx.ofdm_symbol_size, tx.cp_length, tx.data_carriers_per_symbol, tx.bits_per_symbol) == (1024, 128, 511, 1022)
    with pytest.raises(ValueError):
        OFDM.CamG(1024, 32, "16QAM")
    new = OFDM.CamG("A2", "XOR")
    assert not new.old_api and new.K == 2047
