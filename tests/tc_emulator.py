"""TEST DOUBLE: torch-CPU emulation of the tensor-core entry points' documented semantics
(include/rave_b200.h), used ONLY by tests/test_engine_cpu.py to exercise the host-side engine logic
(planning, phase decomposition, pitches, the explicit backward) without a GPU.  Not part of the product."""
import torch
import torch.nn.functional as F


OPERAND_DTYPE = torch.bfloat16


def _bf16(t):
    return t.to(OPERAND_DTYPE)


def conv1d_tc(xa_cl, wt, bias=None, res_cl=None, stride=1, dil=1, pad=(0, 0), act=0, slope=0.2,
              want_f32=True, want_act=False, out_f32=None, out_act=None, out_rows=0, out_row_stride=0,
              out_row_offset=0, Lout=None, res_bf16=None, dact_src=None, Lin=None, res_act=None, res_slope=0.2,
              fm_d=None, fm_partner=None, x3=False, act_cs=0):
    if x3:
        return _conv1d_tc_x3(xa_cl, wt, bias, res_cl, stride, dil, pad, act, slope, out_f32, out_act, out_rows, Lout,
                             Lin, res_act, res_slope, act_cs)
    B, in_pitch, Cin = xa_cl.shape
    Lin = in_pitch if Lin is None else Lin
    K, Cout, _ = wt.shape
    assert in_pitch >= -(-Lin // stride) * stride
    if in_pitch > Lin:
        assert float(xa_cl[:, Lin:].float().abs().max()) == 0.0, "slack rows must be zero"
    x = xa_cl[:, :Lin].float().permute(0, 2, 1)                     # [B, Cin, Lin]
    w = wt.float().permute(1, 2, 0)                                  # [Cout, Cin, K]
    if Lout is None:
        Lout = (Lin + pad[0] + pad[1] - dil * (K - 1) - 1) // stride + 1
    # rows l*stride + k*dil - pad_l, zero outside [0, Lin)
    need = (Lout - 1) * stride + (K - 1) * dil + 1
    pl = pad[0]
    if pl >= 0:
        xp = F.pad(x, (pl, max(0, need - pl - Lin)))
    else:
        xp = F.pad(x[..., -pl:], (0, max(0, need - (Lin + pl))))
    xp = xp[..., :max(need, 1)]
    if xp.shape[-1] < need:
        xp = F.pad(xp, (0, need - xp.shape[-1]))
    y = F.conv1d(xp, w, None, stride, 0, dil)[..., :Lout]            # [B, Cout, Lout]
    v = y.permute(0, 2, 1)                                           # [B, Lout, Cout]
    if bias is not None:
        v = v + bias
    rows = out_rows if out_rows else Lout
    ors = out_row_stride if out_row_stride else 1
    idx = torch.arange(Lout) * ors + out_row_offset
    if dact_src is not None:
        sgn = torch.signbit(dact_src[:, idx].float())
        v = torch.where(sgn, v * slope, v)
    if fm_d is not None and fm_partner is not None:      # fake half only: partner = the real rows
        a = dact_src[:, idx].float()
        ar = fm_partner[:, idx].float()
        hf, hr = torch.where(a > 0, a, a / slope), torch.where(ar > 0, ar, ar / slope)
        v = v - fm_d[0] * torch.sign(hr - hf)
    elif fm_d is not None:
        a = dact_src[:, idx].float()
        h = torch.where(a > 0, a, a / slope)
        hr, hf = h[:B // 2], h[B // 2:]
        sd = torch.sign(hr - hf)
        v = v + torch.cat([fm_d[0] * sd + fm_d[1] * torch.sign(hr), -fm_d[0] * sd], 0)
    if res_bf16 is not None:
        v = v + res_bf16[:, idx].float()
    if res_act is not None:
        ra = res_act[:, idx].float()
        v = v + torch.where(ra > 0, ra, ra / res_slope)
    if res_cl is not None:
        v = v + res_cl[:, idx]
    if want_f32 and out_f32 is None:
        out_f32 = torch.zeros(B, rows, Cout)
    if want_act and out_act is None:
        out_act = torch.zeros(B, rows, Cout, dtype=OPERAND_DTYPE)
    if out_f32 is not None:
        out_f32[:, idx] = v
    if out_act is not None:
        a = F.leaky_relu(v, slope) if act == 1 else v
        out_act[:, idx] = _bf16(a)
    return out_f32, out_act


def _split(v):
    hi = v.to(torch.bfloat16)
    lo = (v - hi.float()).to(torch.bfloat16)
    return hi, lo


def _conv1d_tc_x3(xa_cl, wt, bias, res_cl, stride, dil, pad, act, slope, out_f32, out_act, out_rows, Lout, Lin, res_act,
                  res_slope, act_cs):
    """Split-operand semantics of rave_conv1d_tc_fwd_x3 (include/rave_b200.h): rows [hi | lo], weights [2][K][Cout][Cin],
    hi*hi + lo*hi + hi*lo accumulated in fp32; out_act positions are [hi | lo] pairs of act_cs channels."""
    B, in_pitch, C2 = xa_cl.shape
    Cin = C2 // 2
    K = wt.shape[0] // 2
    Cout = wt.shape[1]
    a_hi, a_lo = xa_cl[..., :Cin].float(), xa_cl[..., Cin:].float()
    w_hi, w_lo = wt[:K].float(), wt[K:].float()
    saved = OPERAND_DTYPE

    def run(a, w):
        o, _ = conv1d_tc(a.to(torch.bfloat16), w.to(torch.bfloat16), None, None, stride, dil, pad, 0, slope,
                         want_f32=True, want_act=False, out_rows=out_rows, Lout=Lout, Lin=Lin)
        return o
    v = run(a_hi, w_hi) + run(a_lo, w_hi) + run(a_hi, w_lo)
    rows = v.shape[1]
    LoutE = Lout if Lout is not None else rows
    if bias is not None:
        v[:, :LoutE] += bias
    if res_act is not None:
        ra = res_act[..., :Cout].float() + res_act[..., Cout:].float()
        v[:, :LoutE] += torch.where(ra > 0, ra, ra / res_slope)[:, :LoutE]
    if res_cl is not None:
        v[:, :LoutE] += res_cl[:, :LoutE]
    if out_f32 is not None:
        out_f32[:, :LoutE] = v[:, :LoutE]
    else:
        out_f32 = None
    if out_act is not None:
        a = F.leaky_relu(v, slope) if act == 1 else v
        cs = act_cs if act_cs else Cout
        hi, lo = _split(a)
        q = Cout // cs
        pair = torch.stack([hi.reshape(B, rows, q, cs), lo.reshape(B, rows, q, cs)], 3)      # [B, rows, q, 2, cs]
        out_act[:, :LoutE] = pair.reshape(B, rows, 2 * Cout)[:, :LoutE]
    return out_f32, out_act


def dilated_unit_tc_supported(C, L):
    return C in (96, 192, 384) and L >= 8


def dilated_unit_tc(xa_cl, w3t, w1t, dil, pad_l, slope_in, slope_mid, act_out, slope_out, L=None, want_a1=False,
                    out_f32=None, out_act=None):
    """Semantics of rave_dilated_unit_tc_fwd: the two per-layer launches back to back, the intermediate rounded to the
    operand type exactly as the kernel rounds it before the second GEMM."""
    B, pitch, C = xa_cl.shape
    L = pitch if L is None else L
    _, a1 = conv1d_tc(xa_cl, w3t, None, None, 1, dil, (pad_l, 2 * dil - pad_l), 1, slope_mid, want_f32=False,
                      want_act=True, out_rows=pitch, Lout=L, Lin=L)
    if pitch > L:
        a1[:, L:] = 0
    conv1d_tc(a1, w1t, None, None, 1, 1, (0, 0), act_out, slope_out, want_f32=False, want_act=False, out_f32=out_f32,
              out_act=out_act, out_rows=pitch, Lout=L, Lin=L, res_act=xa_cl, res_slope=slope_in)
    return (a1 if want_a1 else None), out_f32, out_act


def ncl_to_cl_x3(x):
    xt = x.permute(0, 2, 1).contiguous()
    hi, lo = _split(xt)
    return torch.cat([hi, lo], -1)


def conv1d_tc_wgrad(P_cl, Q_cl, K, stride=1, dil=1, pad_l=0, Lp=None, Lq=None, dbias=None):
    B, p_pitch, Cm = P_cl.shape
    _, q_pitch, Cn = Q_cl.shape
    Lp = p_pitch if Lp is None else Lp
    Lq = q_pitch if Lq is None else Lq
    P = P_cl[:, :Lp].float()
    if dbias is not None:
        dbias += P.sum((0, 1))
    Q = Q_cl[:, :q_pitch].float()
    if q_pitch > Lq:
        assert float(Q[:, Lq:].abs().max()) == 0.0
    S = 2                                  # two "row slices": the batch halves
    dwt = torch.zeros(S, K, Cm, Cn)
    l = torch.arange(Lp)
    half = (B + 1) // 2
    for k in range(K):
        r = l * stride + k * dil - pad_l
        ok = (r >= 0) & (r < Lq)
        if ok.any():
            dwt[0, k] = torch.einsum("blm,bln->mn", P[:half, l[ok]], Q[:half, r[ok]])
            if B > half:
                dwt[1, k] = torch.einsum("blm,bln->mn", P[half:, l[ok]], Q[half:, r[ok]])
    return dwt


def weight_prep_tc(v, g, tapsA, tapsB, C0p, C1p):
    C0, C1 = v.shape[0], v.shape[1]
    v3 = v.reshape(C0, C1, -1)
    norm = None
    w = v3
    if g is not None:
        norm = v3.reshape(C0, -1).norm(2, 1)
        w = v3 * (g.reshape(C0, 1, 1) / norm.reshape(C0, 1, 1))
    wp = F.pad(w, (0, 1, 0, C1p - C1, 0, C0p - C0))          # extra all-zero tap: index -1
    outA = _bf16(wp[:, :, tapsA].permute(2, 0, 1).contiguous()) if tapsA else None
    outB = _bf16(wp[:, :, tapsB].permute(2, 1, 0).contiguous()) if tapsB else None
    return norm, outA, outB


def weight_norm_bwd_tapmajor(dwt, v, g, norm):
    C0, C1 = v.shape[0], v.shape[1]
    dw = dwt.sum(0)[:, :C0, :C1].permute(1, 2, 0).reshape(v.shape)
    if g is None:
        return dw.contiguous(), None
    v2 = v.reshape(C0, -1)
    dw2 = dw.reshape(C0, -1)
    dot = (dw2 * v2).sum(1)
    n = norm
    gg = g.reshape(C0)
    dv = (gg / n).unsqueeze(1) * (dw2 - v2 * (dot / (n * n)).unsqueeze(1))
    dg = (dot / n).reshape(g.shape)
    return dv.reshape(v.shape), dg


def ncl_to_cl(x, act=0, slope=0.2, alpha=None, want_bf16=True, want_f32=False):
    xt = x.permute(0, 2, 1).contiguous()
    a = F.leaky_relu(xt, slope) if act == 1 else xt
    return (_bf16(a) if want_bf16 else None), (xt if want_f32 else None)


def cl_to_ncl(x_cl):
    return x_cl.permute(0, 2, 1).contiguous()


def weight_norm_raw(v, g):
    C0 = v.shape[0]
    norm = v.reshape(C0, -1).norm(2, 1)
    shape = (C0,) + (1,) * (v.dim() - 1)
    return v * (g.reshape(shape) / norm.reshape(shape)), norm


def conv1d_c1(x_rows, w, bias, Lin, stride, pad, act, slope, out_f32=None, out_act=None, Lout=None):
    R, x_pitch = x_rows.shape
    Cout = w.shape[0]
    K = w.numel() // Cout
    x = x_rows[:, :Lin].unsqueeze(1)
    need = (Lout - 1) * stride + K
    xp = F.pad(x, (pad[0], max(0, need - pad[0] - Lin)))
    y = F.conv1d(xp, w.reshape(Cout, 1, K), bias, stride)[..., :Lout].permute(0, 2, 1)
    if out_f32 is not None:
        out_f32[:, :Lout] = y
    if out_act is not None:
        out_act[:, :Lout] = _bf16(F.leaky_relu(y, slope) if act == 1 else y)
    return out_f32, out_act


def conv1d_c1_wgrad(g_cl, x_rows, Cout, K, Lin, Lout, stride, pad_l):
    R = g_cl.shape[0]
    g = g_cl[:, :Lout, :Cout].float()
    dwt = torch.zeros(3, K, Cout, 1)                       # 3 "slices": rows split arbitrarily
    l = torch.arange(Lout)
    for k in range(K):
        pos = l * stride + k - pad_l
        ok = (pos >= 0) & (pos < Lin)
        if ok.any():
            xv = x_rows[:, pos[ok]]                          # [R, n]
            full = torch.einsum("rlc,rl->c", g[:, l[ok]], xv)
            dwt[0, k, :, 0] = 0.25 * full
            dwt[1, k, :, 0] = 0.5 * full
            dwt[2, k, :, 0] = 0.25 * full
    return dwt


def conv1d_c1_dgrad(g_cl, w, x_pitch, Lin, Lout, stride, pad_l):
    R = g_cl.shape[0]
    Cout = w.shape[0]
    K = w.numel() // Cout
    g = g_cl[:, :Lout, :Cout].float().permute(0, 2, 1)               # [R, Cout, Lout]
    full = F.conv_transpose1d(g, w.reshape(Cout, 1, K), None, stride)  # [R, 1, (Lout-1)*s + K]
    dx = torch.zeros(R, x_pitch)
    seg = full[:, 0, pad_l:pad_l + Lin]
    dx[:, :seg.shape[1]] = seg
    return dx


def colsum_bf16(g_cl, L, C):
    return g_cl[:, :L, :C].float().sum((0, 1))


def _c1_rows(src, Lin, period, pool):
    """rows [Bs*period, Lin] derived from src [Bs, T] exactly like the reference: fold (zero pad to a multiple of the
    period) or repeated average pooling."""
    Bs, T = src.shape
    if period > 1:
        xp = F.pad(src, (0, Lin * period - T))
        return xp.reshape(Bs, Lin, period).permute(0, 2, 1).reshape(Bs * period, Lin)
    if pool > 1:
        return src[:, :Lin * pool].reshape(Bs, Lin, pool).mean(-1)
    return src[:, :Lin]


def im2col_c1(src, Lin, Lout, out_pitch, K, stride, pad_l, period=1, pool=1):
    x_rows = _c1_rows(src.float(), Lin, period, pool)
    R = x_rows.shape[0]
    X = torch.zeros(R, out_pitch, 16)
    l = torch.arange(Lout)
    for k in range(K):
        pos = l * stride + k - pad_l
        ok = (pos >= 0) & (pos < Lin)
        X[:, l[ok], k] = x_rows[:, pos[ok]]
    return _bf16(X)


def gather_c1(P_cl, src_shape, Lin, Lout, K, stride, pad_l, period=1, pool=1, batch0=0):
    if batch0:
        part = gather_c1(P_cl, (src_shape[0] - batch0, src_shape[1]), Lin, Lout, K, stride, pad_l, period, pool)
        return torch.cat([torch.zeros(batch0, src_shape[1]), part], 0)
    R = P_cl.shape[0]
    Bs, T = src_shape
    dx = torch.zeros(R, Lin)
    t = torch.arange(Lin)
    for k in range(K):
        q = t + pad_l - k
        ok = (q >= 0) & (q % stride == 0) & (q // stride < Lout)
        dx[:, t[ok]] += P_cl[:, (q[ok] // stride), k]
    # adjoint of _c1_rows
    with torch.enable_grad():
        src = torch.zeros(Bs, T, requires_grad=True)
        rows = _c1_rows(src, Lin, period, pool)
        (g,) = torch.autograd.grad(rows, src, dx)
    return g.detach()


def _unleaky(a, slope):
    return torch.where(a > 0, a, a / slope)


def fm_stats(a_cl, stats_row, L, slope):
    B2 = a_cl.shape[0]
    h = _unleaky(a_cl[:, :L].float(), slope)
    hr, hf = h[:B2 // 2], h[B2 // 2:]
    stats_row[0] += (hr - hf).abs().sum()
    stats_row[1] += hr.abs().sum()


def fm_grad(a_cl, dstats_row, L, slope):
    B2 = a_cl.shape[0]
    h = _unleaky(a_cl.float(), slope)
    hr, hf = h[:B2 // 2], h[B2 // 2:]
    sd = torch.sign(hr - hf)
    g = torch.cat([dstats_row[0] * sd + dstats_row[1] * torch.sign(hr), -dstats_row[0] * sd], 0)
    g[:, L:] = 0
    return _bf16(g)


def score_stats(score_cl, stats6, L):
    B2 = score_cl.shape[0]
    s = score_cl[:, :L, 0].float()
    sr, sf = s[:B2 // 2], s[B2 // 2:]
    stats6.view(-1).add_(torch.stack([(sr - sf).abs().sum(), sr.abs().sum(), torch.relu(1 - sr).sum(),
                           torch.relu(1 + sf).sum(), sr.sum(), sf.sum()]))


def score_grad(score_cl, dstats6, L):
    B2, pitch, C = score_cl.shape
    s = score_cl[:, :L, 0].float()
    sr, sf = s[:B2 // 2], s[B2 // 2:]
    d = dstats6.float().reshape(-1)
    sd = torch.sign(sr - sf)
    gr = d[0] * sd + d[1] * torch.sign(sr) - d[2] * (sr < 1).float() + d[4]
    gf = -d[0] * sd + d[3] * (sf > -1).float() + d[5]
    g = torch.zeros(B2, pitch, C, dtype=torch.float32)
    g[:B2 // 2, :L, 0] = gr
    g[B2 // 2:, :L, 0] = gf
    from rave_b200 import engine
    return g.to(engine.ACT_DTYPE)


def weight_prep_tc_multi(items, x3=False, into=None):
    """`into`: overwrite the (norm, outA, outB) tensors of an earlier call in place (engine.refresh_static_prep)."""
    if into is not None:
        fresh = weight_prep_tc_multi(items, x3=x3)
        for new, old in zip(fresh, into):
            for a, b in zip(new, old):
                if b is not None:
                    b.copy_(a)
        return into
    if not x3:
        return [weight_prep_tc(*it) for it in items]
    out = []
    saved = globals()["OPERAND_DTYPE"]
    globals()["OPERAND_DTYPE"] = torch.float32            # keep fp32, then split into [hi slabs | lo slabs]
    try:
        for it in items:
            norm, A, Bm = weight_prep_tc(*it)
            cat = lambda t: None if t is None else torch.cat(_split(t.float()), 0)
            out.append((norm, cat(A), cat(Bm)))
    finally:
        globals()["OPERAND_DTYPE"] = saved
    return out


def weight_norm_bwd_multi(items):
    out = []
    for it in items:
        dwt, v, g, norm = it[:4]
        if len(it) > 4 and it[4] is not None:          # phase-wide buffer [S][J][C0p][wide*C1p] -> [S][K][C0p][C1p]
            wide, slots = it[4]
            S, J, C0p, W = dwt.shape
            C1p = W // wide
            d5 = dwt.reshape(S, J, C0p, wide, C1p)
            dwt = torch.stack([d5[:, sl // wide, :, sl % wide, :] for sl in slots], 1)
        out.append(weight_norm_bwd_tapmajor(dwt, v, g, norm))
    return out


def snake_cl_fwd(h_cl, alpha):
    al = alpha.detach().reshape(-1).float()
    x = h_cl.float()
    return _bf16(x + torch.sin(al * x) ** 2 / (al + 1e-9))


def snake_cl_bwd(ga_cl, h_cl, alpha, add=None, want_dalpha=True):
    al = alpha.detach().reshape(-1).float()
    ae = al + 1e-9
    x, g = h_cl.float(), ga_cl.float()
    s2 = torch.sin(2 * al * x)
    gh = g * (1 + al * s2 / ae)
    if add is not None:
        gh = gh + add.float()
    dal = (g * (x * s2 / ae - torch.sin(al * x) ** 2 / (ae * ae))).reshape(-1, x.shape[-1]).sum(0) if want_dalpha else None
    return _bf16(gh), dal


def activation(x, act, slope=0.2, alpha=None):
    """ops.activation (fp32 elementwise kernel with its own autograd): plain torch here."""
    if act == 1:
        return F.leaky_relu(x, slope)
    if act == 2:
        al = alpha.reshape(1, -1, *([1] * (x.dim() - 2)))
        return x + torch.sin(al * x) ** 2 / (al + 1e-9)
    return x


def leaky_fm(x, slope):
    """ops.leaky_fm: LeakyReLU + the L1 feature-matching sums of the [real; fake] halves, plain torch autograd here."""
    a = F.leaky_relu(x, slope)
    h = a.shape[0] // 2
    return a, torch.stack([(a[:h] - a[h:]).abs().sum(), a[:h].abs().sum()])


def time_stack_nhwc(x, kt, pt, Cp, Fp):
    """ops.time_stack_nhwc: x [B, T, F, C] fp32 channel-last -> [(b t), Fp, Cp] operand rows holding the kt time-shifted
    copies of the channels side by side (zero outside the T steps, in the pad columns / channels); torch autograd."""
    B, T, Fq, C = x.shape
    xp = F.pad(x, (0, 0, 0, 0, pt, kt - 1 - pt))                             # zero time steps on both sides
    st = torch.cat([xp[:, dt:dt + T] for dt in range(kt)], dim=-1)          # [B, T, F, kt*C], channel = dt*C + c
    st = F.pad(st, (0, Cp - kt * C, 0, Fp - Fq))
    return _bf16(st.reshape(B * T, Fp, Cp))


def leaky_fm_stack(x, slope, T, Fp):
    """ops.leaky_fm_stack = leaky_fm + the next conv's time-stacked operand (kt = 3, pt = 1) from its output."""
    a, st = leaky_fm(x, slope)
    R2, Fq, C = x.shape
    return a, st, time_stack_nhwc(a.view(R2 // T, T, Fq, C), 3, 1, 3 * C, Fp)


def stft_frames(x, window, n_fft, hop):
    """ops.stft_frames: reflect pad (centred STFT) + frame + window, [N, T] -> [N, frames, n_fft]; torch autograd."""
    p = n_fft // 2
    xp = F.pad(x[:, None], (p, p), mode="reflect")[:, 0]
    return xp.unfold(-1, n_fft, hop) * window


def install(monkeypatch):
    from rave_b200 import ops
    for name in ("conv1d_tc", "conv1d_tc_wgrad", "weight_prep_tc", "weight_norm_bwd_tapmajor", "ncl_to_cl",
                 "cl_to_ncl", "weight_norm_raw", "conv1d_c1", "conv1d_c1_wgrad", "fm_stats", "fm_grad", "conv1d_c1_dgrad", "colsum_bf16", "im2col_c1", "gather_c1", "weight_prep_tc_multi",
                 "weight_norm_bwd_multi", "score_stats", "score_grad", "ncl_to_cl_x3", "dilated_unit_tc",
                 "dilated_unit_tc_supported", "snake_cl_fwd", "snake_cl_bwd", "activation", "leaky_fm", "time_stack_nhwc",
                 "leaky_fm_stack", "stft_frames"):
        monkeypatch.setattr(ops, name, globals()[name])
