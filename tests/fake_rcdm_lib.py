"""TEST INFRASTRUCTURE ONLY — a CPU emulation of the handful of C-ABI entry points the stage-1 prior host code calls
(``include/rcdm.h``: rcdm_gemm_ex, rcdm_layernorm, rcdm_masked_attn, rcdm_temporal_attn, rcdm_prior_assemble,
rcdm_unclip_cfg_step, rcdm_pack_geglu, and the folded-LayerNorm calls rcdm_fold_ln / rcdm_gemm_stats_parts /
rcdm_rowstats / rcdm_gemm_ln / rcdm_fold_proj / rcdm_gemm_cat), operating on raw pointers into CPU fp16 tensors.

It exists so that the *host-side orchestration* (argument order, buffer reuse, row pitches, step counter, weight packing
in ``MyPriorTransformer`` / ``Seq_Inpaint_Prior_Pipeline``) can be checked against the oracle in the CPU test tier,
where no GPU is available.  It is never importable from the product: tests monkeypatch ``rcdms_b200._lib.lib`` with
it.  The kernels themselves are checked on the GPU (tests/test_prior_gpu.py).  fp16 only; GEGLU weights keep the
reference layout (``rcdm_pack_geglu`` is the identity here, the epilogue chunks in two)."""
import ctypes

import numpy as np
import torch
import torch.nn.functional as F


def _t(ptr, n, dtype=np.float16):
    if not ptr:
        return None
    nbytes = n * np.dtype(dtype).itemsize
    buf = (ctypes.c_uint8 * nbytes).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=dtype))


def _rn(x):
    return x.half().float()


class FakeLib:
    def __init__(self):
        self.calls = []

    def rcdm_last_error(self):
        return b"fake"

    def rcdm_kernel_launches(self):
        return len(self.calls)

    def rcdm_pack_geglu(self, dt, w, b, wo, bo, N, K, stream):
        _t(wo, N * K).copy_(_t(w, N * K))
        _t(bo, N, np.float32).copy_(_t(b, N, np.float32))
        return 0

    def rcdm_gemm_ex(self, dt, a, lda, w, bias, res, ldr, out, ldo, M, N, K, flags, stream):
        assert dt == 1
        self.calls.append(("gemm", M, N, K, flags))
        A = _t(a, (M - 1) * lda + K).as_strided((M, K), (lda, 1)).float()
        W = _t(w, N * K).reshape(N, K).float()
        y = A @ W.t()
        if bias:
            y = y + _t(bias, N, np.float32)
        n_out = N
        if flags & 1:
            h, g = y.chunk(2, dim=-1)
            y = h * F.gelu(g)
            n_out = N // 2
        elif flags & 2:
            y = F.gelu(y)
        elif flags & 4:
            y = F.silu(y)
        y = _rn(y)
        ldo = ldo if ldo > 0 else n_out
        ldr = ldr if ldr > 0 else N
        if res:
            y = _rn(y + _t(res, (M - 1) * ldr + N).as_strided((M, N), (ldr, 1)).float())
        _t(out, (M - 1) * ldo + n_out).as_strided((M, n_out), (ldo, 1)).copy_(y.half())
        return 0

    # ---- folded LayerNorm (rcdm_fold_ln / rcdm_gemm_stats_parts / rcdm_rowstats / rcdm_gemm_ln / rcdm_fold_proj / rcdm_gemm_cat) ----
    PARTS = 2  # the emulated producer splits its columns into two statistic parts

    def rcdm_gemm_stats_parts(self, M, N):
        return self.PARTS

    def rcdm_fold_ln(self, dt, w, gamma, beta, pe, bias, wf, c, N, K, frames, stream):
        W = _t(w, N * K).reshape(N, K).float()
        g, b = _t(gamma, K, np.float32), _t(beta, K, np.float32)
        wg = W * g
        _t(wf, N * K).reshape(N, K).copy_((wg - wg.mean(dim=1, keepdim=True)).half())
        cc = (W @ b)[None].repeat(frames, 1)
        if pe:
            cc = cc + _t(pe, frames * K, np.float32).reshape(frames, K) @ W.t()
        if bias:
            cc = cc + _t(bias, N, np.float32)
        _t(c, frames * N, np.float32).reshape(frames, N).copy_(cc)
        return 0

    def rcdm_rowstats(self, dt, x, stats, M, K, stream):
        self.calls.append(("rowstats", M, K))
        X = _t(x, M * K).reshape(M, K).float()
        _t(stats, M * 2, np.float32).reshape(M, 2).copy_(torch.stack([X.sum(1), (X * X).sum(1)], dim=1))
        return 0

    def rcdm_gemm_ln(self, dt, a, lda, w, vec, res, out, M, N, K, flags, stats_in, parts_in, frames, rows_per_frame, eps,
                     stats_out, stream):
        assert dt == 1 and lda == K
        self.calls.append(("gemm_ln", M, N, K, flags, bool(stats_in), bool(stats_out)))
        A = _t(a, M * K).reshape(M, K).float()
        y = A @ _t(w, N * K).reshape(N, K).float().t()
        if stats_in:
            st = _t(stats_in, parts_in * M * 2, np.float32).reshape(parts_in, M, 2).double().sum(0)
            mean = st[:, 0] / K
            rstd = (st[:, 1] / K - mean * mean).clamp_min(0).add(eps).rsqrt().float()
            cc = _t(vec, frames * N, np.float32).reshape(frames, N)
            y = y * rstd[:, None] + cc[(torch.arange(M) // max(rows_per_frame, 1)) % frames]
        elif vec:
            y = y + _t(vec, N, np.float32)
        n_out = N
        if flags & 1:
            h, g = y.chunk(2, dim=-1)
            y = h * F.gelu(g)
            n_out = N // 2
        elif flags & 2:
            y = F.gelu(y)
        elif flags & 4:
            y = F.silu(y)
        y = _rn(y)
        if res:
            y = _rn(y + _t(res, M * N).reshape(M, N).float())
        _t(out, M * n_out).reshape(M, n_out).copy_(y.half())
        if stats_out:
            S = _t(stats_out, self.PARTS * M * 2, np.float32).reshape(self.PARTS, M, 2)
            for i, blk in enumerate(y.chunk(self.PARTS, dim=1)):
                S[i, :, 0] = blk.sum(1)
                S[i, :, 1] = (blk * blk).sum(1)
        return 0

    def rcdm_fold_proj(self, dt, wp, w2, b2, bp, wf, cf, C, stream):
        Wp, W2 = _t(wp, C * C).reshape(C, C).float(), _t(w2, C * 4 * C).reshape(C, 4 * C).float()
        _t(wf, C * 5 * C).reshape(C, 5 * C).copy_(torch.cat([Wp, Wp @ W2], dim=1).half())
        _t(cf, C, np.float32).copy_(Wp @ _t(b2, C, np.float32) + _t(bp, C, np.float32))
        return 0

    def rcdm_gemm_cat(self, dt, a0, K0, a1, K1, w, bias, res, out, M, N, stats_out, stream):
        self.calls.append(("gemm_cat", M, N, K0, K1))
        A = torch.cat([_t(a0, M * K0).reshape(M, K0), _t(a1, M * K1).reshape(M, K1)], dim=1).float()
        y = A @ _t(w, N * (K0 + K1)).reshape(N, K0 + K1).float().t()
        if bias:
            y = y + _t(bias, N, np.float32)
        y = _rn(y)
        if res:
            y = _rn(y + _t(res, M * N).reshape(M, N).float())
        _t(out, M * N).reshape(M, N).copy_(y.half())
        if stats_out:
            S = _t(stats_out, self.PARTS * M * 2, np.float32).reshape(self.PARTS, M, 2)
            for i, blk in enumerate(y.chunk(self.PARTS, dim=1)):
                S[i, :, 0] = blk.sum(1)
                S[i, :, 1] = (blk * blk).sum(1)
        return 0

    def rcdm_layernorm(self, dt, x, gamma, beta, out, rows, C, eps, pe, rows_per_frame, frames, stream):
        self.calls.append(("ln", rows, C))
        X = _t(x, rows * C).reshape(rows, C).float()
        y = F.layer_norm(X, (C,), _t(gamma, C, np.float32), _t(beta, C, np.float32), eps)
        if pe:
            P = _t(pe, frames * C, np.float32).reshape(frames, C)
            y = y + P[(torch.arange(rows) // rows_per_frame) % frames]
        _t(out, rows * C).reshape(rows, C).copy_(y.half())
        return 0

    def rcdm_masked_attn(self, dt, qkv, ld, key_bias, causal, out, ldo, batch, heads, S, d, stream):
        self.calls.append(("mattn", batch, heads, S, d))
        C = heads * d
        Q = _t(qkv, batch * S * ld).reshape(batch, S, ld).float()
        q, k, v = [Q[..., i * C:(i + 1) * C].reshape(batch, S, heads, d).transpose(1, 2) for i in range(3)]
        s = q @ k.transpose(-1, -2) * d ** -0.5
        if key_bias:
            s = s + _t(key_bias, batch * S, np.float32).reshape(batch, 1, 1, S)
        if causal:
            s = s + torch.full((S, S), -10000.0).triu_(1)
        o = (_rn(s.softmax(-1)) @ v).transpose(1, 2).reshape(batch, S, C)
        _t(out, batch * S * ldo).reshape(batch, S, ldo)[..., :C].copy_(o.half())
        return 0

    def rcdm_temporal_attn(self, dt, qkv, out, batch, frames, hw, heads, d, stream):
        self.calls.append(("tattn", batch, frames, hw, heads, d))
        C = heads * d
        Q = _t(qkv, batch * frames * hw * 3 * C).reshape(batch, frames, hw, 3 * C).float()
        q, k, v = [Q[..., i * C:(i + 1) * C].reshape(batch, frames, hw, heads, d).permute(0, 2, 3, 1, 4)
                   for i in range(3)]
        o = (q @ k.transpose(-1, -2) * d ** -0.5).softmax(-1) @ v
        _t(out, batch * frames * hw * C).reshape(batch, frames, hw, C).copy_(
            o.permute(0, 3, 1, 2, 4).reshape(batch, frames, hw, C).half())
        return 0

    def rcdm_prior_assemble(self, dt, base, temb, hproj, pos, x, B, S, C, t_row, h_row, n_lat, step, stream):
        self.calls.append(("assemble", B, S, C, t_row, h_row, n_lat))
        st = int(_t(step, 1, np.int32)[0]) if step else 0
        X = _t(x, B * S * C).reshape(B, S, C)
        X.copy_(_t(base, B * S * C).reshape(B, S, C))
        P = _t(pos, S * C).reshape(S, C)
        X[:, t_row] = _t(temb, (st + 1) * C).reshape(-1, C)[st] + P[t_row]
        X[:, h_row] = _t(hproj, n_lat * C).reshape(n_lat, C)[torch.arange(B) % n_lat] + P[h_row]
        return 0

    def rcdm_unclip_cfg_step(self, dt, pred, lat, noise, coef, n, do_cfg, g, step, advance, stream):
        self.calls.append(("unclip", n, do_cfg))
        st_t = _t(step, 1, np.int32)
        st = int(st_t[0])
        c = _t(coef, (st + 1) * 8, np.float32).reshape(-1, 8)[st]
        P = _t(pred, (2 if do_cfg else 1) * n)
        p = P[:n] + g * (P[n:] - P[:n]) if do_cfg else P
        x = _t(lat, n)
        x0 = p if c[6] == 0 else (x - float(c[3]) * p) / float(c[4])
        if c[5] > 0:
            x0 = x0.clamp(-float(c[5]), float(c[5]))
        new = float(c[0]) * x0 + float(c[1]) * x
        if c[2] > 0:
            new = new + float(c[2]) * _t(noise, (st + 1) * n).reshape(-1, n)[st]
        x.copy_(new)
        if advance:
            st_t[0] = st + 1
        return 0
