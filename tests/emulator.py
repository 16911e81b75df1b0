"""CPU emulator of the native op table (TEST INFRASTRUCTURE).

Executes the planner IR (audioldm2_b200/plan.py) op by op with torch on byte buffers that stand
in for the device arena / workspace.  Semantics follow include/aldm_b200.h; arithmetic is plain
fp32 (operand planes are emulated as fp16 hi [+ lo] exactly like the kernels split them; a missing lo plane
-- ``None`` -- is a single-plane operand), so a
run checks graph wiring, weight layout / K ordering, epilogue flags and the plan-time buffer
allocator (a liveness bug corrupts the result) -- everything except the CUDA kernels themselves.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from audioldm2_b200 import _lib, packing
from audioldm2_b200.plan import Plan, Ref


class Emulator:
    def __init__(self, plan: Plan):
        self.plan = plan
        self.mem = {"w": plan.arena.clone(), "ws": torch.zeros(plan.ws_bytes + 1024, dtype=torch.uint8)}
        # poison the workspace so that reads of never-written memory are visible
        self.mem["ws"].view(torch.float32)[:] = float("nan")

    # ---- typed views ---------------------------------------------------------------------
    def f32(self, ref: Ref, n: int) -> torch.Tensor:
        return self.mem[ref.region][ref.off:ref.off + 4 * n].view(torch.float32)

    def f16(self, ref: Ref, n: int) -> torch.Tensor:
        return self.mem[ref.region][ref.off:ref.off + 2 * n].view(torch.float16)

    def i64(self, ref: Ref, n: int) -> torch.Tensor:
        return self.mem[ref.region][ref.off:ref.off + 8 * n].view(torch.int64)

    def write_io(self, name: str, value: torch.Tensor):
        kind, ref, shape = self.plan.io[name]
        n = int(np.prod(shape))
        assert tuple(value.shape) == tuple(shape), (name, value.shape, shape)
        if kind == "i64":
            self.i64(ref, n)[:] = value.reshape(-1).to(torch.int64)
        else:
            self.f32(ref, n)[:] = value.reshape(-1).float()

    def read_io(self, name: str) -> torch.Tensor:
        kind, ref, shape = self.plan.io[name]
        return self.f32(ref, int(np.prod(shape))).reshape(shape).clone()

    def write_planes(self, hi: Ref, lo, x: torch.Tensor, ld: int, rows: int, col0: int = 0):
        h, l = packing.split_f16(x)
        H = self.f16(hi, rows * ld).reshape(rows, ld)
        H[:, col0:col0 + x.shape[1]] = h
        if lo is not None:
            L = self.f16(lo, rows * ld).reshape(rows, ld)
            L[:, col0:col0 + x.shape[1]] = l

    def read_planes(self, hi: Ref, lo, rows: int, ld: int) -> torch.Tensor:
        x = self.f16(hi, rows * ld).reshape(rows, ld).float()
        return x if lo is None else x + self.f16(lo, rows * ld).reshape(rows, ld).float()

    def strided_planes(self, hi: Ref, lo, rows: int, cols: int, ld: int):
        """(hi, lo-or-None) fp16 views [rows, cols] with row stride ld"""
        n = (rows - 1) * ld + cols
        H = torch.as_strided(self.f16(hi, n), (rows, cols), (ld, 1))
        L = None if lo is None else torch.as_strided(self.f16(lo, n), (rows, cols), (ld, 1))
        return H, L

    def put_planes(self, hi: Ref, lo, rows: int, cols: int, ld: int, idx, v: torch.Tensor):
        h, l = packing.split_f16(v)
        H, L = self.strided_planes(hi, lo, rows, cols, ld)
        H[idx] = h
        if L is not None:
            L[idx] = l

    # ---- ops -----------------------------------------------------------------------------
    def run(self, first: int = 0, last=None):
        for o in self.plan.ops[first:last]:
            getattr(self, "op_" + o["kind"])(o)

    def op_copy(self, o):
        n = o["bytes"]
        self.mem[o["dst"].region][o["dst"].off:o["dst"].off + n] = self.mem[o["src"].region][o["src"].off:o["src"].off + n].clone()

    def op_temb(self, o):
        B, dim = o["B"], o["dim"]
        t = self.i64(o["t"], B).float()
        freqs = self.f32(o["freqs"], dim // 2)
        args = t[:, None] * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], -1)
        self.write_planes(o["out_hi"], o.get("out_lo"), emb, dim, B)

    def op_prep(self, o):
        rows, c0, c1, Cp, mode = o["rows"], o["c0"], o["c1"], o["Cp"], o["mode"]
        if o["src_nchw"]:
            HW = o["HW"]; B = rows // HW
            x = self.f32(o["src0"], rows * c0).reshape(B, c0, HW).permute(0, 2, 1).reshape(rows, c0)
        else:
            x = self.f32(o["src0"], rows * c0).reshape(rows, c0)
        if c1:
            x = torch.cat([x, self.f32(o["src1"], rows * c1).reshape(rows, c1)], 1)
        x = x.clone()
        assert torch.isfinite(x).all(), "prep reads uninitialised / non-finite memory"
        Cc = c0 + c1
        if mode in (_lib.PREP_GN, _lib.PREP_GN_SILU):
            B, HW = o["B"], o["HW"]
            g = self.f32(o["gamma"], Cc); b = self.f32(o["beta"], Cc)
            y = F.group_norm(x.reshape(B, HW, Cc).permute(0, 2, 1), o["groups"], g, b, o["eps"]).permute(0, 2, 1).reshape(rows, Cc)
            if mode == _lib.PREP_GN_SILU:
                y = F.silu(y)
        elif mode == _lib.PREP_LN:
            y = F.layer_norm(x, (Cc,), self.f32(o["gamma"], Cc), self.f32(o["beta"], Cc), o["eps"])
        elif mode == _lib.PREP_SILU:
            y = F.silu(x)
        elif mode == _lib.PREP_LRELU:
            y = F.leaky_relu(x, o["slope"])
        else:
            y = x
        if Cp > Cc:
            y = torch.cat([y, torch.zeros(rows, Cp - Cc)], 1)
        self.write_planes(o["out_hi"], o.get("out_lo"), y, Cp, rows)

    def op_packb(self, o):
        N, K, bn, lds = o["N"], o["K"], o["bn"], o["lds"]
        if o["transpose"]:
            src = self.f32(o["src"], K * lds).reshape(K, lds)[:, :N].t()
        else:
            src = self.f32(o["src"], N * lds).reshape(N, lds)[:, :K]
        packed, plain, Npad, Kpad = packing.pack_tiles(src.clone().contiguous(), bn)
        self.mem[o["dst_packed"].region][o["dst_packed"].off:o["dst_packed"].off + packed.numel()] = packed
        if o.get("dst_plain") is not None:
            self.f32(o["dst_plain"], Npad * Kpad)[:] = plain.reshape(-1)

    def op_softmax(self, o):
        rows, n = o["rows"], o["n"]
        x = self.f32(o["x"], rows * n).reshape(rows, n) * o["scale"]
        self.write_planes(o["out_hi"], o.get("out_lo"), F.softmax(x, -1), n, rows)

    def op_attn(self, o):
        B, h, Nq, Nk = o["B"], o["heads"], o["Nq"], o["Nk"]
        d = 32
        Cc = h * d
        bmod = o["kv_bmod"]
        Bkv = bmod if bmod > 0 else B

        def planes(hi, rows, ld, col):       # the attention op reads the hi plane only (single-plane fp16 Q, K, V)
            n = (rows - 1) * ld + col + Cc
            return torch.as_strided(self.f16(hi, n).float(), (rows, Cc), (ld, 1), col)

        q = planes(o["q_hi"], B * Nq, o["ldq"], o["q_col"]).reshape(B, Nq, h, d).permute(0, 2, 1, 3)
        k = planes(o["k_hi"], Bkv * Nk, o["ldk"], o["k_col"]).reshape(Bkv, Nk, h, d).permute(0, 2, 1, 3)
        ld_t = o["ld_t"]
        nvt = Bkv * Cc * ld_t
        vt = self.f16(o["vt_hi"], nvt).float().reshape(Bkv, h, d, ld_t)[..., :Nk]
        v = vt.permute(0, 1, 3, 2)                                  # [Bkv, h, Nk, d]
        if bmod > 0:
            idx = torch.arange(B) % bmod
            k, v = k[idx], v[idx]
        assert torch.isfinite(q).all() and torch.isfinite(k).all() and torch.isfinite(v).all(), "attention reads garbage"
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * o["scale"]
        if o.get("mask") is not None:
            m = self.f32(o["mask"], Bkv * Nk).reshape(Bkv, 1, 1, Nk)
            if bmod > 0:
                m = m[torch.arange(B) % bmod]
            sim = sim.masked_fill(~(m == 1), -torch.finfo(torch.float32).max)
        # the kernel rounds the un-normalised probabilities exp(s - max) to fp16 before P V and divides by the fp32 row sum
        p = torch.exp(sim - sim.amax(-1, keepdim=True))
        out = torch.einsum("bhij,bhjd->bhid", p.half().float(), v) / p.sum(-1, keepdim=True)
        out = out.permute(0, 2, 1, 3).reshape(B * Nq, Cc)
        self.write_planes(o["out_hi"], o.get("out_lo"), out, o["ldo"], B * Nq)

    def op_gemm(self, o):
        B, H, W, Cp, up = o["B"], o["H"], o["W"], o["Cp"], o["up"]
        if o["splitk"] > 1:
            # the kernel scribbles partial sums over its split-K workspace: emulate the clobber so an
            # overlap with a live buffer is caught
            n = o["splitk"] * packing.round_up(B * o["OH"] * o["OW"], 128) * packing.round_up(o["N"], o["bn"])
            self.f32(o["ws"], n)[:] = float("nan")
        OH, OW, sy, sx = o["OH"], o["OW"], o["sy"], o["sx"]
        Hs, Ws = H >> up, W >> up
        bmod = o["bmod"]
        Bsrc = bmod if bmod > 0 else B
        A = self.read_planes(o["a_hi"], o.get("a_lo"), Bsrc * Hs * Ws, Cp).reshape(Bsrc, Hs, Ws, Cp)
        assert torch.isfinite(A).all(), "gemm reads uninitialised / non-finite operand planes"
        if bmod > 0:
            A = A[torch.arange(B) % bmod]
        if up:
            A = A.repeat_interleave(2, 1).repeat_interleave(2, 2)
        N, K, Kpad, bn = o["N"], o["K"], o["Kpad"], o["bn"]
        Npad = packing.round_up(N, bn)
        if o.get("w_plain") is not None:
            Wm = self.f32(o["w_plain"], Npad * Kpad).reshape(Npad, Kpad)
        else:
            nbytes = Npad * Kpad * 4
            Wm = packing.unpack_tiles(self.mem[o["w_packed"].region][o["w_packed"].off:o["w_packed"].off + nbytes], Npad, Kpad, bn)
        M = B * OH * OW
        acc = torch.zeros(M, Npad)
        oh = torch.arange(OH)[:, None]
        ow = torch.arange(OW)[None, :]
        for t, (dy, dx) in enumerate(o["taps"]):
            ih, iw = oh * sy + dy, ow * sx + dx
            valid = ((ih >= 0) & (ih < H) & (iw >= 0) & (iw < W)).expand(OH, OW)
            ihc, iwc = ih.clamp(0, H - 1).expand(OH, OW), iw.clamp(0, W - 1).expand(OH, OW)
            g = A[:, ihc, iwc, :] * valid[None, :, :, None]
            acc += g.reshape(M, Cp) @ Wm[:, t * Cp:(t + 1) * Cp].t()
        b_idx = torch.arange(M) // (OH * OW)
        if o.get("bias") is not None:
            acc = acc + self.f32(o["bias"], Npad)[None]
        if o.get("rowvec") is not None:
            ld = o["ld_rowvec"]
            total = (B - 1) * ld + Npad
            rv = torch.as_strided(self.f32(o["rowvec"], total), (B, N), (ld, 1))
            acc[:, :N] += rv[b_idx]
        act = o["act"]
        if act == _lib.ACT_GEGLU:
            t = acc.reshape(M, Npad // bn, 2, bn // 2)
            v = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(M, Npad // 2)
            n_out = N // 2
        else:
            v = torch.tanh(acc) if act == _lib.ACT_TANH else (F.silu(acc) if act == _lib.ACT_SILU else acc)
            n_out = N
        v = v[:, :n_out].clone()
        # output row mapping
        ohh = (torch.arange(M) // OW) % OH
        oww = torch.arange(M) % OW
        orow = (b_idx * o["OHF"] + ohh * o["osy"] + o["ooy"]) * o["OWF"] + oww
        nrows_out = B * o["OHF"] * o["OWF"]
        if o.get("res") is not None:
            ld = o["ld_res"]
            R = torch.as_strided(self.f32(o["res"], (nrows_out - 1) * ld + n_out), (nrows_out, n_out), (ld, 1))
            assert torch.isfinite(R[orow]).all(), "gemm residual reads uninitialised memory"
            v = v + R[orow]
        v = v * o["alpha"]
        mode = o["out_mode"]
        if mode == _lib.OUT_F32:
            ld = o["ldo"]
            O = torch.as_strided(self.f32(o["out"], (nrows_out - 1) * ld + n_out), (nrows_out, n_out), (ld, 1))
            if o["accumulate"]:
                v = v + O[orow]
            O[orow] = v
            if o.get("out_hi") is not None:      # dual output (fp32 + operand planes)
                self.put_planes(o["out_hi"], o.get("out_lo"), nrows_out, n_out, ld, orow, v)
        elif mode == _lib.OUT_QKV:
            ns, tpb, ld_t, ld = o["n_split"], o["tok_per_batch"], o["ld_t"], o["ldo"]
            self.put_planes(o["out_hi"], o.get("out_lo"), M, ns, ld, slice(None), v[:, :ns])
            Cv, nb = N - ns, M // tpb
            t = v[:, ns:].reshape(nb, tpb, Cv).permute(0, 2, 1)              # [b, c, tok]
            tp = torch.zeros(nb, Cv, ld_t); tp[:, :, :tpb] = t
            th, tl = packing.split_f16(tp)
            self.f16(o["out2_hi"], nb * Cv * ld_t)[:] = th.reshape(-1)
            if o.get("out2_lo") is not None:
                self.f16(o["out2_lo"], nb * Cv * ld_t)[:] = tl.reshape(-1)
        elif mode == _lib.OUT_PLANES:
            self.put_planes(o["out_hi"], o.get("out_lo"), nrows_out, n_out, o["ldo"], orow, v)
        else:   # NCHW
            O = self.f32(o["out"], B * N * OH * OW).reshape(B, N, OH, OW)
            O[:] = v.reshape(B, OH, OW, N).permute(0, 3, 1, 2)
