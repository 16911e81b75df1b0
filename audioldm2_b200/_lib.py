"""ctypes binding of the C-ABI in include/aldm_b200.h (+ the in-tree nvcc build).

The shared library is built IN-TREE (audioldm2_b200/libaldm_b200.so) so it travels with the
repo snapshot to the GPU box.  There is no fallback: if the library is missing or the device
is not sm_100, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_PATH = os.path.join(HERE, "libaldm_b200.so")
SOURCES = ["gemm.cu", "prep.cu", "attention.cu", "elementwise.cu", "stft.cu", "program.cu", "engine_abi.cu", "microbench.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math=false"]

MAX_TAPS = 16
ABI_VERSION = 6

# enums (keep in sync with the header; checked by tests/test_abi.py against the header text)
GEMM_TC, GEMM_SIMT, GEMM_TC_V1 = 0, 1, 2
GEMM_STATIC_B = 1 << 16
ACT_NONE, ACT_GEGLU, ACT_TANH, ACT_SILU = 0, 1, 2, 3
OUT_F32, OUT_PLANES, OUT_NCHW, OUT_QKV = 0, 1, 2, 3
PREP_COPY, PREP_SILU, PREP_LRELU, PREP_GN, PREP_GN_SILU, PREP_LN = 0, 1, 2, 3, 4, 5
OP_GEMM, OP_PREP, OP_ATTN, OP_SOFTMAX, OP_TEMB, OP_TRANSPOSE, OP_PACKB, OP_COPY = 1, 2, 3, 4, 5, 6, 7, 8


class GemmDesc(C.Structure):
    _fields_ = [
        ("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("w_packed", C.c_void_p), ("w_plain", C.c_void_p),
        ("bias", C.c_void_p), ("rowvec", C.c_void_p), ("res", C.c_void_p), ("out", C.c_void_p),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("out2_hi", C.c_void_p), ("out2_lo", C.c_void_p),
        ("ws", C.c_void_p),
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cp", C.c_int32),
        ("up", C.c_int32), ("bmod", C.c_int32),
        ("OH", C.c_int32), ("OW", C.c_int32), ("sy", C.c_int32), ("sx", C.c_int32),
        ("ntaps", C.c_int32),
        ("dy", C.c_int16 * MAX_TAPS), ("dx", C.c_int16 * MAX_TAPS),
        ("N", C.c_int32), ("K", C.c_int32), ("Kpad", C.c_int32), ("bn", C.c_int32),
        ("ldo", C.c_int32), ("ld_res", C.c_int32), ("ld_rowvec", C.c_int32),
        ("OHF", C.c_int32), ("OWF", C.c_int32), ("osy", C.c_int32), ("ooy", C.c_int32),
        ("act", C.c_int32), ("out_mode", C.c_int32), ("accumulate", C.c_int32), ("splitk", C.c_int32),
        ("impl", C.c_int32), ("n_split", C.c_int32), ("tok_per_batch", C.c_int32), ("ld_t", C.c_int32),
        ("alpha", C.c_float),
    ]


class PrepDesc(C.Structure):
    _fields_ = [
        ("src0", C.c_void_p), ("src1", C.c_void_p), ("gamma", C.c_void_p), ("beta", C.c_void_p),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("scratch", C.c_void_p),
        ("rows", C.c_int32), ("c0", C.c_int32), ("c1", C.c_int32), ("Cp", C.c_int32),
        ("B", C.c_int32), ("HW", C.c_int32), ("groups", C.c_int32), ("mode", C.c_int32),
        ("eps", C.c_float), ("slope", C.c_float), ("src_nchw", C.c_int32),
    ]


class AttnDesc(C.Structure):
    _fields_ = [
        ("q_hi", C.c_void_p), ("q_lo", C.c_void_p), ("k_hi", C.c_void_p), ("k_lo", C.c_void_p),
        ("vt_hi", C.c_void_p), ("vt_lo", C.c_void_p), ("mask", C.c_void_p),
        ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
        ("B", C.c_int32), ("heads", C.c_int32), ("Nq", C.c_int32), ("Nk", C.c_int32),
        ("ldq", C.c_int32), ("ldk", C.c_int32), ("ld_t", C.c_int32), ("ldo", C.c_int32),
        ("q_col", C.c_int32), ("k_col", C.c_int32), ("kv_bmod", C.c_int32), ("impl", C.c_int32),
        ("scale", C.c_float),
    ]


class _Softmax(C.Structure):
    _fields_ = [("x", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
                ("rows", C.c_int32), ("n", C.c_int32), ("scale", C.c_float)]


class _Temb(C.Structure):
    _fields_ = [("t", C.c_void_p), ("freqs", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
                ("B", C.c_int32), ("dim", C.c_int32)]


class _Transpose(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("B", C.c_int32), ("C", C.c_int32),
                ("HW", C.c_int32), ("to_nhwc", C.c_int32)]


class _PackB(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst_packed", C.c_void_p), ("dst_plain", C.c_void_p),
                ("lds", C.c_int32), ("transpose", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("bn", C.c_int32)]


class _Copy(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("bytes", C.c_int64)]


class _OpU(C.Union):
    _fields_ = [("gemm", GemmDesc), ("prep", PrepDesc), ("attn", AttnDesc), ("softmax", _Softmax),
                ("temb", _Temb), ("transpose", _Transpose), ("packb", _PackB), ("copy", _Copy)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("tag", C.c_int32), ("u", _OpU)]


MAX_LANES = 8


class UnetLane(C.Structure):
    """aldm_unet_lane: one independent sub-batch of the UNet (own programs + workspace slots)."""
    _fields_ = [("cond", C.c_void_p), ("step", C.c_void_p), ("x_slot", C.c_void_p), ("t_slot", C.c_void_p),
                ("eps_slot", C.c_void_p), ("ctx_slot", C.c_void_p * 2), ("mask_slot", C.c_void_p * 2),
                ("film_slot", C.c_void_p)]


class EngineDesc(C.Structure):
    """aldm_engine_desc (include/aldm_b200.h): programs + their fixed I/O slots."""
    _fields_ = [("lane", UnetLane * MAX_LANES), ("n_lanes", C.c_int32),
                ("vae_dec", C.c_void_p), ("vocoder", C.c_void_p), ("vae_enc", C.c_void_p),
                ("z_slot", C.c_void_p), ("mel_slot", C.c_void_p), ("voc_mel_slot", C.c_void_p), ("wave_slot", C.c_void_p),
                ("enc_mel_slot", C.c_void_p), ("moments_slot", C.c_void_p), ("B", C.c_int32), ("latent_elems", C.c_int32),
                ("mel_elems", C.c_int32), ("wave_len", C.c_int32), ("n_ctx", C.c_int32), ("ctx_len", C.c_int32 * 2),
                ("ctx_dim", C.c_int32 * 2), ("film_dim", C.c_int32), ("use_graph", C.c_int32)]


def build(verbose: bool = False, force: bool = False) -> str:
    """Compile every CUDA source for sm_100a into audioldm2_b200/libaldm_b200.so (nvcc cross-compiles
    without a GPU).  Rebuilds only when a source is newer than the library."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "common.cuh"), os.path.join(ROOT, "include", "aldm_b200.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in srcs:
        o = os.path.join(HERE, "build", os.path.basename(s) + ".o")
        objs.append(o)
        cmd = ["nvcc"] + [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")] + ["-c", s, "-o", o]
        if verbose:
            print(" ".join(cmd))
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode()))
        if verbose and out:
            print(out.decode())
    cmd = ["nvcc", "-shared", "-o", LIB_PATH] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(cmd), r.stdout.decode()))
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the shared library (raises if it has not been built -- no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`. "
                           "There is no CPU/PyTorch fallback for the native path.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sig = {
        "aldm_gemm": (i32, [C.POINTER(GemmDesc), vp]),
        "aldm_prep": (i32, [C.POINTER(PrepDesc), vp]),
        "aldm_pack_b": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, vp]),
        "aldm_attention": (i32, [C.POINTER(AttnDesc), vp]),
        "aldm_softmax_rows": (i32, [vp, i32, i32, f32, vp, vp, vp]),
        "aldm_timestep_embedding": (i32, [vp, i32, i32, vp, vp, vp, vp]),
        "aldm_ddim_step": (i32, [vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, vp]),
        "aldm_masked_blend": (i32, [vp, vp, vp, vp, i32, i32, i32, f32, f32, vp]),
        "aldm_transpose_chw": (i32, [vp, vp, i32, i32, i32, i32, vp]),
        "aldm_posterior_sample": (i32, [vp, vp, vp, i32, i32, i32, f32, vp]),
        "aldm_stft_mel": (i32, [vp, i32, i32, i32, i32, vp, i32, vp, i32, vp]),
        "aldm_program_create": (i32, [C.POINTER(Op), i32, C.POINTER(vp)]),
        "aldm_program_run": (i32, [vp, vp]),
        "aldm_program_run_range": (i32, [vp, i32, i32, vp]),
        "aldm_program_capture": (i32, [vp, vp]),
        "aldm_program_replay": (i32, [vp, vp]),
        "aldm_program_num_launches": (i32, [vp]),
        "aldm_program_is_captured": (i32, [vp]),
        "aldm_program_destroy": (None, [vp]),
        "aldm_engine_create": (i32, [C.POINTER(EngineDesc), C.POINTER(vp)]),
        "aldm_engine_destroy": (None, [vp]),
        "aldm_engine_set_conditioning": (i32, [vp, i32, vp, vp, i32, vp, vp, i32, vp, vp]),
        "aldm_engine_precompute": (i32, [vp, vp]),
        "aldm_engine_unet_eps": (i32, [vp, vp, i64, vp, vp, vp]),
        "aldm_engine_ddim_step": (i32, [vp, vp, i64, vp, f32, f32, f32, f32, f32, vp, vp, vp]),
        "aldm_engine_vae_decode": (i32, [vp, vp, vp, vp]),
        "aldm_engine_vocoder": (i32, [vp, vp, vp, vp]),
        "aldm_engine_vae_encode": (i32, [vp, vp, vp, vp]),
        "aldm_sizeof_engine_desc": (C.c_size_t, []),
        "aldm_abi_version": (i32, []),
        "aldm_sizeof_op": (C.c_size_t, []),
        "aldm_sizeof_gemm_desc": (C.c_size_t, []),
        "aldm_offsetof_gemm": (C.c_size_t, [i32]),
        "aldm_last_error": (C.c_char_p, []),
        "aldm_device_check": (i32, [i32]),
        "aldm_debug_timeline": (i32, [vp, i32]),
        "aldm_debug_umma_rate": (i32, [i32, i32, i32, vp, i32]),
        "aldm_debug_store_rate": (i32, [i32, i32, i32, i64, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if L.aldm_abi_version() != ABI_VERSION:
        raise RuntimeError("libaldm_b200.so ABI version mismatch: rebuild")
    if L.aldm_sizeof_engine_desc() != C.sizeof(EngineDesc):
        raise RuntimeError("ctypes mirror of aldm_engine_desc does not match the C layout")
    if L.aldm_sizeof_op() != C.sizeof(Op) or L.aldm_sizeof_gemm_desc() != C.sizeof(GemmDesc):
        raise RuntimeError("ctypes mirror of aldm_op / aldm_gemm_desc does not match the C layout")
    _lib = L
    return L


EXPORTED = ["aldm_gemm", "aldm_prep", "aldm_pack_b", "aldm_attention", "aldm_softmax_rows",
            "aldm_timestep_embedding", "aldm_ddim_step", "aldm_masked_blend", "aldm_transpose_chw",
            "aldm_posterior_sample", "aldm_stft_mel", "aldm_program_create", "aldm_program_run",
            "aldm_program_run_range", "aldm_program_capture", "aldm_program_replay",
            "aldm_program_num_launches", "aldm_program_is_captured", "aldm_program_destroy", "aldm_engine_create",
            "aldm_engine_destroy", "aldm_engine_set_conditioning", "aldm_engine_precompute", "aldm_engine_unet_eps",
            "aldm_engine_ddim_step", "aldm_engine_vae_decode", "aldm_engine_vocoder", "aldm_engine_vae_encode",
            "aldm_sizeof_engine_desc", "aldm_abi_version", "aldm_sizeof_op",
            "aldm_sizeof_gemm_desc", "aldm_offsetof_gemm", "aldm_last_error", "aldm_device_check", "aldm_debug_timeline",
            "aldm_debug_umma_rate", "aldm_debug_store_rate"]


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().aldm_last_error().decode(errors="replace")
        raise RuntimeError(f"aldm error {rc} {what}: {msg}")


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
