"""CPU: the C-ABI library loads without a GPU and exports every function include/mrssm_b200.h declares; the ctypes
prototype table binds only declared functions.  No compute call is made."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mrssm_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mrssm_[a-z0-9_]+)\s*\(", src)))


def _lib():
    from mrssm_b200 import _lib as L
    if not os.path.exists(L.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "multimodal-rssm_b200", "csrc")], stdout=subprocess.DEVNULL)
    return L, ctypes.CDLL(L.LIB_PATH)


def test_library_exports_every_declared_function():
    L, lib = _lib()
    names = _declared()
    assert len(names) > 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_table_matches_header():
    L, lib = _lib()
    names = set(_declared())
    unknown = [n for n in L.SYMBOLS if n not in names]
    assert not unknown, unknown
    L.load()                                    # declares every prototype; raises on a missing symbol
    assert L.load().mrssm_abi_version() >= 1


def test_rollout_tc_plan_is_host_only():
    """The tcgen05 rollout's MMA program is built on the host: sizes of the shipped configs are eligible, the plan fits
    its tables, larger models are rejected (they run the CUDA-core kernels)."""
    L, lib = _lib()
    L.load()
    pb, kb = ctypes.c_int64(), ctypes.c_int64()
    for E in (0, 1, 3):
        assert lib.mrssm_rollout_tc_eligible(200, 30, 200, 3, E) == 1
        L.call_host("mrssm_rollout_tc_plan_bytes", 200, 30, 200, 3, E, ctypes.byref(pb), ctypes.byref(kb))
        buf = ctypes.create_string_buffer(pb.value)
        L.call_host("mrssm_rollout_tc_plan", 200, 30, 200, 3, E, ctypes.cast(buf, ctypes.c_void_p), pb.value)
        hdr = (ctypes.c_int32 * 13).from_buffer(buf)
        D, S, H, A, NH, n_tiles, n_ops, n_pack = hdr[1:9]
        assert (D, S, H, A, NH) == (200, 30, 200, 3, 1 + E)
        assert n_ops == n_pack and 0 < n_tiles <= 112 and n_ops <= 336
        # bytes of one step's weight stream: every MMA carries a [2][N][8] bf16 block
        assert kb.value % 512 == 0 and kb.value > 0
    assert lib.mrssm_rollout_tc_eligible(1024, 64, 1024, 3, 3) == 0      # BASELINE config 5: CUDA-core rollout
    assert lib.mrssm_rollout_tc_eligible(200, 30, 200, 3, 4) == 0
