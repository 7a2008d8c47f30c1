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


STRUCTS = {"mrssm_t4": "T4", "mrssm_conv_args": "ConvArgs", "mrssm_tc_conv_args": "TcConvArgs", "mrssm_tv": "TV",
           "mrssm_pl_conv_args": "PlConvArgs", "mrssm_rollout_args": "RolloutArgs", "mrssm_rollout_bwd_args": "RolloutBwdArgs",
           "mrssm_latent_args": "LatentArgs", "mrssm_overshoot_args": "OvershootArgs",
           "mrssm_replay_gather_args": "ReplayGatherArgs", "mrssm_gconv_args": "GConvArgs", "mrssm_norm_args": "NormArgs",
           "mrssm_rstep_ws": "RstepWs"}


def test_header_is_plain_c_and_struct_layouts_match_ctypes(tmp_path):
    """The boundary is a C ABI: the header compiles as C99 (no C++ / torch types), every struct it defines has a ctypes
    mirror, and size plus the offset of every field agree between gcc and ctypes (a reordered or mistyped field in either
    place would silently shift every pointer behind it)."""
    from mrssm_b200 import _lib as L
    src = open(HEADER).read()
    assert sorted(re.findall(r"^} (mrssm_[a-z0-9_]+);", src, flags=re.M)) == sorted(STRUCTS)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "%s"' % HEADER, 'int main(void) {']
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for field in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, field[0], cname, field[0]))
    lines += ['return 0; }']
    c_file, exe = tmp_path / "layout.c", tmp_path / "layout"
    c_file.write_text("\n".join(lines))
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(c_file)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, pyname in STRUCTS.items():
        cls = getattr(L, pyname)
        assert int(got[cname]) == ctypes.sizeof(cls), (cname, got[cname], ctypes.sizeof(cls))
        for field in cls._fields_:
            assert int(got[f"{cname}.{field[0]}"]) == getattr(cls, field[0]).offset, (cname, field[0])
