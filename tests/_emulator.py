"""TEST-ONLY descriptor emulators.

``torch_b200.ops`` funnels every kernel launch through two functions (``_execute_contract`` and ``_execute_conv``).
Here they are replaced by CPU interpreters of the *same descriptors* the CUDA kernels receive, so that the host logic
(index classification, sub-dim merging, stride bookkeeping, spec strings of every functional path and of their autograd
formulas) is verified on the GPU-less build box.  Nothing in the product imports this file.
"""
import contextlib
import math

import torch
import torch.nn.functional as F

from torch_b200 import ops

LOG = []


def _view(t, sizes, strides):
    return torch.as_strided(t, list(sizes), list(strides), storage_offset=t.storage_offset())


def emulate_contract(plan, a, b, d, out, alpha):
    if plan.swap:
        a, b = b, a
    f = plan.desc_fields
    sb, sm, sn, sk = f["size_b"], f["size_m"], f["size_n"], f["size_k"]
    for name in ("b", "m", "n", "k"):
        assert len(f["size_" + name]) <= 4
    A = _view(a, sb + sm + sk, f["a_b"] + f["a_m"] + f["a_k"])
    B = _view(b, sb + sk + sn, f["b_b"] + f["b_k"] + f["b_n"])
    nb, nm, nn, nk = len(sb), len(sm), len(sn), len(sk)
    L = "abcdefghijklmnopqrstuvwxyz"
    lb, lm, ln, lk = L[:nb], L[nb:nb + nm], L[nb + nm:nb + nm + nn], L[nb + nm + nn:nb + nm + nn + nk]
    res = torch.einsum(f"{lb}{lm}{lk},{lb}{lk}{ln}->{lb}{lm}{ln}", A, B) * alpha
    if d is not None:
        res = res + _view(d, sb + sm + sn, f["d_b"] + f["d_m"] + f["d_n"])
    _view(out, sb + sm + sn, f["c_b"] + f["c_m"] + f["c_n"]).copy_(res)
    LOG.append(("contract", plan.spec, dict(plan.classes), plan.tc is not None))


def _geom(desc):
    n = desc.ndim
    g = dict(ndim=n, batch=desc.batch, channels=desc.channels)
    for k in ("in_size", "out_size", "kernel", "stride", "padding", "dilation"):
        g[k] = [int(getattr(desc, k)[i]) for i in range(n)]
    return g


def _tap_slices(g, taps):
    sl = [slice(None), slice(None)]
    for i in range(g["ndim"]):
        start = taps[i] * g["dilation"][i]
        sl.append(slice(start, start + (g["out_size"][i] - 1) * g["stride"][i] + 1, g["stride"][i]))
    return tuple(sl)


def _pad_arg(g):
    pad = []
    for p in reversed(g["padding"]):
        pad += [p, p]
    return pad


def _unpad(xp, g):
    sl = [slice(None), slice(None)]
    for i in range(g["ndim"]):
        p = g["padding"][i]
        sl.append(slice(p, p + g["in_size"][i]))
    return xp[tuple(sl)]


def _all_taps(g):
    import itertools
    return itertools.product(*[range(k) for k in g["kernel"]])


def _im2col(x, g):
    n = g["ndim"]
    B, C = x.shape[0], x.shape[1]
    xp = F.pad(x, _pad_arg(g))
    cols = x.new_zeros((B, *g["out_size"], C, *g["kernel"]))
    for taps in _all_taps(g):
        patch = xp[_tap_slices(g, taps)]                         # (B, C, *out)
        cols[(slice(None),) + (slice(None),) * n + (slice(None),) + taps] = patch.movedim(1, -1)
    return cols.reshape(B * math.prod(g["out_size"]), C * math.prod(g["kernel"]))


def _col2im(dcols, g, B, C):
    n = g["ndim"]
    padded = [s + 2 * p for s, p in zip(g["in_size"], g["padding"])]
    dxp = dcols.new_zeros((B, C, *padded))
    cols = dcols.reshape(B, *g["out_size"], C, *g["kernel"])
    for taps in _all_taps(g):
        part = cols[(slice(None),) + (slice(None),) * n + (slice(None),) + taps]      # (B, *out, C)
        dxp[_tap_slices(g, taps)] += part.movedim(-1, 1)
    return _unpad(dxp, g)


def emulate_conv(which, desc, *tensors):
    g = _geom(desc)
    n = g["ndim"]
    conv = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[n]
    if which == "tlt_im2col":
        x, cols = tensors
        cols.copy_(_im2col(x, g))
    elif which == "tlt_col2im":
        dcols, dx = tensors
        dx.copy_(_col2im(dcols, g, dx.shape[0], dx.shape[1]))
    elif which == "tlt_dwconv_fwd":
        x, w, y = tensors
        y.copy_(conv(x, w.reshape(w.shape[0], 1, *g["kernel"]), stride=g["stride"], padding=g["padding"],
                     dilation=g["dilation"], groups=x.shape[1]))
    elif which == "tlt_dwconv_dgrad":
        dy, w, dx = tensors
        B, C = dx.shape[0], dx.shape[1]
        padded = [s + 2 * p for s, p in zip(g["in_size"], g["padding"])]
        dxp = dy.new_zeros((B, C, *padded))
        wk = w.reshape(C, *g["kernel"])
        for taps in _all_taps(g):
            dxp[_tap_slices(g, taps)] += dy * wk[(slice(None),) + taps].reshape(1, C, *([1] * n))
        dx.copy_(_unpad(dxp, g))
    elif which == "tlt_dwconv_wgrad":
        x, dy, dw = tensors
        xp = F.pad(x, _pad_arg(g))
        dwk = dw.reshape(x.shape[1], *g["kernel"])
        for taps in _all_taps(g):
            dwk[(slice(None),) + taps] += (xp[_tap_slices(g, taps)] * dy).sum(dim=[0] + list(range(2, 2 + n))).to(dw.dtype)
    else:
        raise AssertionError(which)
    LOG.append((which, g))


def emulate_btt3_reconstruct(c1, c2, c3, W, rows, cols, r1, r2):
    assert list(c1.shape) == [1, rows[0], cols[0], r1] and list(c2.shape) == [r1, rows[1], cols[1], r2]
    assert list(c3.shape) == [r2, rows[2], cols[2], 1]
    W.copy_(torch.einsum("aoib,bpjc,cqkd->opqijk", c1, c2, c3).reshape(W.shape))
    LOG.append(("blocktt3_reconstruct", tuple(rows), tuple(cols), r1, r2))


def emulate_btt3_core_grads(c1, c2, c3, dW32, d1, d2, d3, rows, cols, r1, r2):
    # autograd is unavailable inside a custom-op body: write the three chain-rule contractions out
    g6 = dW32.double().reshape(*rows, *cols)                                   # (o, p, q, i, j, k)
    C1, C2, C3 = c1.double(), c2.double(), c3.double()
    g = (torch.einsum("opqijk,bpjc,cqkd->oib", g6, C2, C3).unsqueeze(0),
         torch.einsum("opqijk,aoib,cqkd->bpjc", g6, C1, C3),
         torch.einsum("opqijk,aoib,bpjc->cqk", g6, C1, C2).unsqueeze(-1))
    for dst, src in zip((d1, d2, d3), g):
        dst.copy_(src.to(dst.dtype))
    LOG.append(("blocktt3_core_grads", tuple(rows), tuple(cols), r1, r2))


def emulate_gemm(a, b, bias, out, M, N, K, lda, ldb, a_major, b_major, alpha=1.0, split_k=0):
    A = torch.as_strided(a, (M, K), (1, lda) if a_major else (lda, 1), a.storage_offset())
    Bm = torch.as_strided(b, (N, K), (1, ldb) if b_major else (ldb, 1), b.storage_offset())
    res = alpha * (A.double() @ Bm.double().t())
    if bias is not None:
        res = res + bias.double()
    torch.as_strided(out, (M, N), (out.stride(0), 1), out.storage_offset()).copy_(res.to(out.dtype))   # ldc = out.stride(0)
    LOG.append(("gemm", M, N, K, a_major, b_major))


def emulate_pointwise(x3, w_nk, bias, y):
    res = torch.einsum("nk,bks->bns", w_nk.double(), x3.double())
    if bias is not None:
        res = res + bias.double()[None, :, None]
    y.copy_(res.to(y.dtype))
    LOG.append(("pointwise_tc", tuple(x3.shape), tuple(w_nk.shape), tuple(w_nk.stride())))


def emulate_pointwise_wgrad(gy3, x3, dw32):
    K = x3.shape[1]
    dw32[:, :K].copy_(torch.einsum("bns,bks->nk", gy3.double(), x3.double()).to(dw32.dtype))
    LOG.append(("pointwise_wgrad_tc", tuple(gy3.shape), tuple(x3.shape)))


def emulate_mmd(x4, mats, flags, addend, y4):
    res = x4.double()
    for k in range(3):
        if mats[k] is None:
            continue
        m = mats[k].double()
        m = m.t() if flags[k] else m                      # (r, d)
        res = torch.movedim(torch.tensordot(res, m, dims=([1 + k], [1])), -1, 1 + k)
    if addend is not None:
        res = res + addend.double()
    y4.copy_(res.to(y4.dtype))
    LOG.append(("multi_mode_dot", tuple(x4.shape), tuple(y4.shape), list(flags)))


def _mmd16_chain(x, dims_in, mats, flags):
    res = x
    for k in range(3):
        if mats[k] is None:
            continue
        m = mats[k].double()
        m = m.t() if flags[k] else m                      # (r, d)
        res = torch.movedim(torch.tensordot(res, m, dims=([1 + k], [1])), -1, 1 + k)
    return res


def emulate_mmd16_fwd(x2, dims_in, dims_out, mats, flags, addend, y2):
    B, n_in, n_out = x2.shape[0], math.prod(dims_in), math.prod(dims_out)
    assert x2.stride(0) % 8 == 0 and y2.stride(0) % 8 == 0 and max(*dims_in, *dims_out) <= 16
    res = _mmd16_chain(x2[:, :n_in].double().reshape(B, *dims_in), dims_in, mats, flags).reshape(B, n_out)
    if addend is not None:
        res = res + addend.double().reshape(-1)
    y2.zero_()
    y2[:, :n_out].copy_(res.to(y2.dtype))
    LOG.append(("mmd16_fwd", tuple(dims_in), tuple(dims_out), list(flags)))


def emulate_mmd16_bwd(u2, g2, dims_in, dims_out, mats, flags, du2, dM):
    B, n_in, n_out = g2.shape[0], math.prod(dims_in), math.prod(dims_out)
    g = g2[:, :n_out].double().reshape(B, *dims_out)
    if du2 is not None:
        gu = _mmd16_chain(g, dims_out, mats, [not f for f in flags])          # multiply by the transposed matrices
        du2.zero_()
        du2[:, :n_in].copy_(gu.reshape(B, n_in).to(du2.dtype))
    if dM is not None:
        u = u2[:, :n_in].double().reshape(B, *dims_in)
        letters = "pqr"
        for k in range(3):
            if mats[k] is None:
                continue
            others = [m if j != k else None for j, m in enumerate(mats)]
            z = _mmd16_chain(u, dims_in, others, flags)                       # every mode but k projected
            gm = torch.einsum("b" + letters.replace(letters[k], "j") + ",b" + letters.replace(letters[k], "i") + "->ji", g, z)
            dM[k, :dims_out[k], :dims_in[k]] += gm.to(dM.dtype)
    LOG.append(("mmd16_bwd", tuple(dims_in), tuple(dims_out), list(flags)))


def emulate_pack_kmajor(src2, dst):
    dst.zero_()
    dst[:, :src2.shape[1]].copy_(src2.to(dst.dtype))
    LOG.append(("pack_kmajor", tuple(src2.shape), dst.stride(0)))


def emulate_nchw_to_rows(x3, rows):
    b, c, sp = x3.shape
    rows.zero_()
    rows[:, :c].copy_(x3.permute(0, 2, 1).reshape(b * sp, c))
    LOG.append(("nchw_to_rows", tuple(x3.shape), rows.stride(0)))


def emulate_rows_to_nchw(rows, y3):
    b, c, sp = y3.shape
    y3.copy_(rows[:, :c].reshape(b, sp, c).permute(0, 2, 1))
    LOG.append(("rows_to_nchw", tuple(y3.shape), rows.stride(0)))


def emulate_unfold_rows(z, cols, batch, in_size, out_size, kernel, stride, padding, dilation, fold):
    nd, ld = len(in_size), z.shape[1]
    taps = math.prod(kernel)
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]
    # unfold as a conv with a one-hot kernel per tap: (taps, 1, *kernel) applied per channel
    eye = torch.eye(taps, dtype=torch.float64).reshape(taps, 1, *kernel)

    def unfold(zz):
        img = zz.double().reshape(batch, *in_size, ld).movedim(-1, 1).reshape(batch * ld, 1, *in_size)
        out = fn(img, eye, None, stride, padding, dilation)                       # (B ld, taps, *out)
        out = out.reshape(batch, ld, taps, math.prod(out_size)).permute(0, 3, 2, 1)
        return out.reshape(batch * math.prod(out_size), taps * ld)

    if not fold:
        cols.copy_(unfold(z).to(cols.dtype))
    else:
        with torch.enable_grad():
            zz = torch.zeros(z.shape, dtype=torch.float64)
            # fold = transpose of unfold: use the linear map explicitly
            basis = torch.func.vjp(unfold, zz)[1](cols.double())[0]
        z.copy_(basis.to(z.dtype))
    LOG.append(("col2im_rows" if fold else "im2col_rows", tuple(in_size), tuple(kernel), tuple(stride), tuple(padding), tuple(dilation)))


def emulate_skinny_fwd(a2, k2, y2):
    y2.copy_((a2.double() @ k2.double().t()).to(y2.dtype))
    LOG.append(("skinny_fwd", tuple(a2.shape), tuple(k2.shape)))


def emulate_skinny_bwd(a2, g2, k2, da2, dk32):
    if da2 is not None:
        da2.copy_((g2.double() @ k2.double()).to(da2.dtype))
    if dk32 is not None:
        dk32 += (g2.double().t() @ a2.double()).to(dk32.dtype)
    LOG.append(("skinny_bwd", tuple(g2.shape), tuple(k2.shape)))


def emulate_dwconv_rows(kind, a, b, out, batch, in_size, out_size, kernel, stride, padding, dilation):
    nd, ld = len(in_size), a.shape[1]
    fn = {1: F.conv1d, 2: F.conv2d, 3: F.conv3d}[nd]

    def img(rows, size):
        return rows.double().reshape(batch, *size, ld).movedim(-1, 1)

    def rows_of(t):
        return t.movedim(1, -1).reshape(-1, ld)

    if kind == "fwd":
        w = b.double().t().reshape(ld, 1, *kernel)
        out.copy_(rows_of(fn(img(a, in_size), w, None, stride, padding, dilation, ld)).to(out.dtype))
    else:
        with torch.enable_grad():
            if kind == "dgrad":
                zz = torch.zeros((batch, ld, *in_size), dtype=torch.float64)
                w = b.double().t().reshape(ld, 1, *kernel)
                f = lambda t: fn(t, w, None, stride, padding, dilation, ld)
                res = torch.func.vjp(f, zz)[1](img(a, out_size).contiguous())[0]
                out.copy_(rows_of(res).to(out.dtype))
            else:
                zz = img(a, in_size).contiguous()
                w0 = torch.zeros((ld, 1, *kernel), dtype=torch.float64)
                f = lambda w: fn(zz, w, None, stride, padding, dilation, ld)
                res = torch.func.vjp(f, w0)[1](img(b, out_size).contiguous())[0]
                out += res.reshape(ld, -1).t().to(out.dtype)
    LOG.append(("dwconv_rows_" + kind, tuple(in_size), tuple(kernel), tuple(stride), tuple(padding), tuple(dilation)))


def emulate_dwsep_taps(kh, kw, lam, ld, flip, taps):
    r = kh.shape[1]
    t = torch.cat([kh.double(), kw.double() * lam.double()], 0)                       # (6, R)
    if flip:
        t = torch.cat([t[:3].flip(0), t[3:].flip(0)], 0)
    taps.zero_()
    taps[:, :r] = t.to(taps.dtype)
    LOG.append(("dwsep_taps", r, ld, bool(flip)))


def emulate_dwsep(kind, a, b, out, batch, height, width, stride):
    ld = a.shape[1]
    ho, wo = height // stride, width // stride

    def img(rows, h, w):
        return rows.double().reshape(batch, h, w, ld).movedim(-1, 1).contiguous()

    def rows_of(t):
        return t.movedim(1, -1).reshape(-1, ld)

    def weight(taps):
        t = taps.double()
        return torch.einsum("ir,jr->rij", t[:3], t[3:]).reshape(ld, 1, 3, 3)

    conv = lambda t, w: F.conv2d(t, w, None, stride, 1, 1, ld)      # noqa: E731
    if kind == "fwd":
        out.copy_(rows_of(conv(img(a, height, width), weight(b))).to(out.dtype))
    elif kind == "dgrad":
        taps = torch.cat([b[:3].flip(0), b[3:].flip(0)], 0) if stride == 1 else b     # stride 1 hands the reversed taps
        w = weight(taps)
        with torch.enable_grad():
            zz = torch.zeros((batch, ld, height, width), dtype=torch.float64)
            res = torch.func.vjp(lambda t: conv(t, w), zz)[1](img(a, ho, wo))[0]
        out.copy_(rows_of(res).to(out.dtype))
    else:
        zz = img(a, height, width)
        with torch.enable_grad():
            w0 = torch.zeros((ld, 1, 3, 3), dtype=torch.float64)
            res = torch.func.vjp(lambda w: conv(zz, w), w0)[1](img(b, ho, wo))[0]
        out.zero_()
        out[0].copy_(res.reshape(ld, 9).t().to(out.dtype))
    LOG.append(("dwsep_rows_" + kind, (height, width), stride))


def emulate_dwsep_finalize(acc, kh, kw, lam, g_kh, g_kw, g_lam):
    r = kh.shape[1]
    t = acc.double().sum(0)[:, :r].reshape(3, 3, r)                                   # [i][j][r]
    h, w, l = kh.double(), kw.double(), lam.double()
    g_kh.copy_((l * torch.einsum("jr,ijr->ir", w, t)).to(g_kh.dtype))
    gwl = torch.einsum("ir,ijr->jr", h, t)
    g_kw.copy_((l * gwl).to(g_kw.dtype))
    g_lam.copy_((w * gwl).sum(0).to(g_lam.dtype))
    LOG.append(("dwsep_finalize", r))


def emulate_chain_prep(c_in, c_out, rank, flip_d, f_in, f_out, kh, kw, lam, wp_in, wp_out, taps_f, taps_d):
    wp_in.zero_(); wp_out.zero_()
    wp_in[:rank, :c_in] = f_in.t()
    wp_out[:c_out, :rank] = f_out
    emulate_dwsep_taps(kh, kw, lam, taps_f.shape[1], False, taps_f)
    emulate_dwsep_taps(kh, kw, lam, taps_d.shape[1], flip_d, taps_d)
    del LOG[-2:]
    LOG.append(("cpchain_prep", c_in, c_out, rank))


def emulate_chain_post(c_in, c_out, rank, dw_out, dw_in, acc, kh, kw, lam, g_fout, g_fin, g_kh, g_kw, g_lam):
    g_fout.copy_(dw_out[:, :rank].to(g_fout.dtype))
    g_fin.copy_(dw_in[:, :rank].to(g_fin.dtype))
    emulate_dwsep_finalize(acc, kh, kw, lam, g_kh, g_kw, g_lam)
    del LOG[-1:]
    LOG.append(("cpchain_post", c_in, c_out, rank))


def _kr_ref(factors, weights):
    R = factors[0].shape[1]
    res = torch.ones((1, R), dtype=torch.float64) if weights is None else weights.double().reshape(1, R)
    for f in factors:
        res = (res[:, None, :] * f.double()[None, :, :]).reshape(-1, R)
    return res


def emulate_kr_fwd(factors, weights, out):
    out.copy_(_kr_ref(factors, weights).to(out.dtype))
    LOG.append(("khatri_rao_fwd", [tuple(f.shape) for f in factors], weights is not None))


def emulate_kr_bwd(factors, weights, g, dfs, dw):
    R = factors[0].shape[1]
    dims = [f.shape[0] for f in factors]
    gt = g.double().reshape(*dims, R)
    letters = "abcd"[:len(factors)]
    w = torch.ones(R, dtype=torch.float64) if weights is None else weights.double()
    for k in range(len(factors)):
        if dfs[k] is None:
            continue
        others = [factors[j].double() for j in range(len(factors)) if j != k]
        spec = letters + "r," + ",".join(letters[j] + "r" for j in range(len(factors)) if j != k) + "->" + letters[k] + "r"
        t = torch.einsum(spec, gt, *others) if others else gt
        dfs[k].copy_((t * w).to(dfs[k].dtype))
        if k == 0 and dw is not None:
            dw += (t * factors[0].double()).sum(0).to(dw.dtype)
    LOG.append(("khatri_rao_bwd", [tuple(f.shape) for f in factors], weights is not None))


def _kr_full(fs):
    """(prod dims, R) Khatri-Rao matrix, first factor slowest, fp64."""
    out = fs[0].double()
    for f in fs[1:]:
        out = (out[:, None, :] * f.double()[None, :, :]).reshape(-1, f.shape[1])
    return out


def emulate_cplinear_fwd(d, x, f_in, f_out, lam, bias, z, y):
    """tlt_cplinear_fwd: the descriptor's dims must describe the factors it is given; z keeps round4(R) columns, pad = 0."""
    assert [int(d.in_dims[k]) for k in range(d.n_in)] == [f.shape[0] for f in f_in]
    assert [int(d.out_dims[k]) for k in range(d.n_out)] == [f.shape[0] for f in f_out]
    R = int(d.rank)
    zz = x.double() @ _kr_full(f_in)
    z.zero_()
    z[:, :R] = zz.to(z.dtype)
    yy = (zz * lam.double()) @ _kr_full(f_out).t()
    if bias is not None:
        yy = yy + bias.double()
    y.copy_(yy.to(y.dtype))
    LOG.append(("cplinear_fwd", tuple(x.shape), R))


def emulate_cplinear_bwd(d, x, gy, z, f_in, f_out, lam, g, dx, df_in, df_out, dlam, dbias):
    R = int(d.rank)
    kin, kout = _kr_full(f_in), _kr_full(f_out)
    G = gy.double() @ kout
    zz = z[:, :R].double()
    dx.copy_(((G * lam.double()) @ kin.t()).to(dx.dtype))
    dlam.copy_((zz * G).sum(0).to(dlam.dtype))
    if dbias is not None:
        dbias.copy_(gy.double().sum(0).to(dbias.dtype))

    def factor_grads(dkr, fs, outs):
        dims = [f.shape[0] for f in fs]
        t = dkr.reshape(*dims, R)
        letters = "abcd"[:len(fs)]
        for k in range(len(fs)):
            others = [fs[j].double() for j in range(len(fs)) if j != k]
            spec = letters + "r" + "".join("," + letters[j] + "r" for j in range(len(fs)) if j != k) + "->" + letters[k] + "r"
            outs[k].copy_((torch.einsum(spec, t, *others) if others else t).to(outs[k].dtype))

    factor_grads(x.double().t() @ (G * lam.double()), f_in, df_in)
    factor_grads(gy.double().t() @ (zz * lam.double()), f_out, df_out)
    LOG.append(("cplinear_bwd", tuple(x.shape), R))


def emulate_split3(src, rows, cols, ld, pat_cols, out_cols, pat_rows, out_rows):
    """tlt_split3_bf16: hi / mid / lo bf16 pieces in the six-block patterns (A: h h m h m l, B: h m h l m h), pad columns zero."""
    x = torch.as_strided(src, (rows, cols), (ld, 1)).float()
    hi = x.to(torch.bfloat16); r = x - hi.float()
    mid = r.to(torch.bfloat16); r = r - mid.float()
    lo = r.to(torch.bfloat16)
    pieces = [hi, mid, lo]
    pats = {0: [0, 0, 1, 0, 1, 2], 1: [0, 1, 0, 2, 1, 0]}
    cp = (cols + 7) // 8 * 8
    if out_cols is not None:
        out_cols.zero_()
        for b, q in enumerate(pats[pat_cols]):
            out_cols[:, b * cp:b * cp + cols] = pieces[q]
    if out_rows is not None:
        out_rows.zero_()
        for b, q in enumerate(pats[pat_rows]):
            out_rows[b * rows:(b + 1) * rows, :cols] = pieces[q]
    LOG.append(("split3", rows, cols))


def emulate_kr_split(factors, weights, pat_cols, out_cols, pat_rows, out_rows):
    kr = _kr_ref(factors, weights).float().contiguous()
    emulate_split3(kr, kr.shape[0], kr.shape[1], kr.shape[1], pat_cols, out_cols, pat_rows, out_rows)


def emulate_kr_bwd_ld(factors, weights, g, ldg, dfs, dw):
    R = factors[0].shape[1]
    emulate_kr_bwd(factors, weights, torch.as_strided(g, (g.shape[0], R), (ldg, 1)).contiguous(), dfs, dw)


def emulate_kron_fwd(a, b, k):
    k.copy_(torch.einsum("ap,bq->abpq", a.double(), b.double()).reshape(k.shape).to(k.dtype))
    LOG.append(("kron2_fwd", tuple(a.shape), tuple(b.shape)))


def emulate_kron_bwd(a, b, dk, d_a, d_b):
    g = dk.double().reshape(a.shape[0], b.shape[0], a.shape[1], b.shape[1])
    d_a.copy_(torch.einsum("abpq,bq->ap", g, b.double()).to(d_a.dtype))
    d_b.copy_(torch.einsum("abpq,ap->bq", g, a.double()).to(d_b.dtype))
    LOG.append(("kron2_bwd", tuple(a.shape), tuple(b.shape)))


def emulate_trl_tail_fwd(d, zp, core, fo, bias, y, v):
    B, J, rc, ldz, ro, O = (int(getattr(d, k)) for k in ("batch", "j", "rc", "ldz", "ro", "out"))
    assert tuple(zp.shape) == (B, J, ldz) and core.numel() == rc * J * ro and tuple(fo.shape) == (O, ro)
    vv = torch.einsum("bjc,cjr->br", zp[:, :, :rc].double(), core.reshape(rc, J, ro).double())
    v.copy_(vv.to(v.dtype))
    yy = vv @ fo.double().t()
    if bias is not None:
        yy = yy + bias.double()
    y.copy_(yy.to(y.dtype))
    LOG.append(("trl_tail_fwd", (B, J, rc, ro, O)))


def emulate_trl_tail_bwd(d, zp, gy, core, fo, v, dv, dzp, dcore, dfo, dbias, phases=3):
    B, J, rc, ldz, ro, O = (int(getattr(d, k)) for k in ("batch", "j", "rc", "ldz", "ro", "out"))
    g = gy.double()
    if phases & 1:
        dvv = g @ fo.double()
        dv.copy_(dvv.to(dv.dtype))
        dzp.zero_()
        dzp[:, :, :rc] = torch.einsum("br,cjr->bjc", dvv, core.reshape(rc, J, ro).double()).to(dzp.dtype)
    if phases & 2:                                   # reads the dv the first phase wrote (fp32), like the kernel
        dvv = dv.double()
        dcore.copy_(torch.einsum("bjc,br->cjr", zp[:, :, :rc].double(), dvv).reshape(dcore.shape).to(dcore.dtype))
        dfo.copy_((g.t() @ v.double()).to(dfo.dtype))
        if dbias is not None:
            dbias.copy_(g.sum(0).to(dbias.dtype))
    LOG.append(("trl_tail_bwd", (B, J, rc, ro, O), phases))


def emulate_scale_modes(x, masks, alpha, y):
    res = x.double() * alpha
    for k, m in enumerate(masks):
        if m is not None:
            shape = [1] * x.dim()
            shape[k] = -1
            res = res * m.double().reshape(shape)
    y.copy_(res.to(y.dtype))
    LOG.append(("scale_modes", tuple(x.shape), [m is not None for m in masks], alpha))


def emulate_spatial_fwd(x3, k, t):
    t.copy_(torch.einsum("bcs,sj->bjc", x3.double(), k.double()).to(t.dtype))
    LOG.append(("spatial_project_fwd", tuple(x3.shape), tuple(k.shape)))


def emulate_spatial_bwd(x3, k, gt, dx, dk):
    dx.copy_(torch.einsum("bjc,sj->bcs", gt.double(), k.double()).to(dx.dtype))
    dk.add_(torch.einsum("bcs,bjc->sj", x3.double(), gt.double()).to(dk.dtype))
    LOG.append(("spatial_project_bwd", tuple(x3.shape), tuple(k.shape)))


_CPROWS_WS, _CPROWS_ACC = {}, {}     # workspace / accumulator address -> what the kernels would have put there


def _cprows_chain(x, lam, f_out, f_in, kh, kw, stride):
    """(B, H, W, C) rows -> (B, H/s, W/s, O): 1x1 -> depthwise (3, 1) -> depthwise (1, 3) -> 1x1, in the given precision."""
    r = lam.shape[0]
    z = torch.einsum("bhwc,cr->brhw", x, f_in)
    z = F.conv2d(z, kh.t().reshape(r, 1, 3, 1), stride=(stride, 1), padding=(1, 0), groups=r)
    z = F.conv2d(z, (kw * lam).t().reshape(r, 1, 1, 3), stride=(1, stride), padding=(0, 1), groups=r)
    return torch.einsum("brhw,or->bhwo", z, f_out)


def emulate_cprows_pack(d, f_in, f_out, kh, kw, lam, ws):
    _CPROWS_WS[ws.data_ptr()] = tuple(t.detach().clone() for t in (lam, f_out, f_in, kh, kw))
    LOG.append(("cpconv_rows_pack", int(d.in_channels), int(d.out_channels), int(d.rank)))


def emulate_cprows_fwd(d, x, ws, bias, y):
    lam, f_out, f_in, kh, kw = (t.double() for t in _CPROWS_WS[ws.data_ptr()])
    out = _cprows_chain(x.double(), lam, f_out, f_in, kh, kw, int(d.stride))
    if bias is not None:
        out = out + bias.double()
    y.copy_(out.to(y.dtype))
    LOG.append(("cpconv_rows_fwd", tuple(x.shape), int(d.rank), int(d.stride), bias is not None))


def emulate_cprows_bwd(d, x, gy, ws, dx, acc):
    """Closed-form backward of _cprows_chain (custom-op bodies run below autograd, so no torch.autograd.grad here)."""
    from torch.nn.grad import conv2d_input, conv2d_weight
    lam, f_out, f_in, kh, kw = (t.double() for t in _CPROWS_WS[ws.data_ptr()])
    st, r = int(d.stride), lam.shape[0]
    x64, g64 = x.double(), gy.double()
    wh, ww = kh.t().reshape(r, 1, 3, 1).contiguous(), (kw * lam).t().reshape(r, 1, 1, 3).contiguous()
    z1 = torch.einsum("bhwc,cr->brhw", x64, f_in)
    u = F.conv2d(z1, wh, stride=(st, 1), padding=(1, 0), groups=r)
    z2 = F.conv2d(u, ww, stride=(1, st), padding=(0, 1), groups=r)
    dz2 = torch.einsum("bhwo,or->brhw", g64, f_out).contiguous()
    g_fout = torch.einsum("bhwo,brhw->or", g64, z2)
    du = conv2d_input(u.shape, ww, dz2, stride=(1, st), padding=(0, 1), groups=r)
    g_kwl = conv2d_weight(u, ww.shape, dz2, stride=(1, st), padding=(0, 1), groups=r).reshape(r, 3).t()
    dz1 = conv2d_input(z1.shape, wh, du, stride=(st, 1), padding=(1, 0), groups=r)
    g_kh = conv2d_weight(z1, wh.shape, du, stride=(st, 1), padding=(1, 0), groups=r).reshape(r, 3).t()
    dx.copy_(torch.einsum("brhw,cr->bhwc", dz1, f_in).to(dx.dtype))
    g_fin = torch.einsum("bhwc,brhw->cr", x64, dz1)
    _CPROWS_ACC[acc.data_ptr()] = ((kw * g_kwl).sum(0), g_fout, g_fin, g_kh, g_kwl * lam)
    LOG.append(("cpconv_rows_bwd", tuple(x.shape), int(d.rank)))


def emulate_cprows_finalize(d, acc, kh, kw, lam, g_fout, g_fin, g_kh, g_kw, g_lam):
    gl, gfo, gfi, gkh, gkw = _CPROWS_ACC.pop(acc.data_ptr())
    for dst, src in ((g_lam, gl), (g_fout, gfo), (g_fin, gfi), (g_kh, gkh), (g_kw, gkw)):
        dst.copy_(src.to(dst.dtype))
    LOG.append(("cpconv_rows_finalize", int(d.rank)))


@contextlib.contextmanager
def emulated_kernels():
    """Swap the launch funnels for the CPU interpreters (tests only)."""
    from torch_b200 import blocktt_ops as bo
    from torch_b200 import pointwise_ops as po
    from torch_b200 import modedot_ops as mo
    from torch_b200 import tucker_ops as to
    from torch_b200 import rows_ops as ro
    from torch_b200 import kr_ops as ko
    from torch_b200 import spatial_ops as so
    from torch_b200 import cprows_ops as co
    from torch_b200 import cplinear_ops as cl
    from torch_b200 import cplinear_tc_ops as ct
    saved_ct = ct._execute_split, ct._execute_kr_bwd_ld, ct._execute_kr_split
    ct._execute_split, ct._execute_kr_bwd_ld, ct._execute_kr_split = emulate_split3, emulate_kr_bwd_ld, emulate_kr_split
    from torch_b200 import trl_ops as tr
    saved_tr = tr._execute_fwd, tr._execute_bwd, tr._execute_kron_fwd, tr._execute_kron_bwd
    tr._execute_fwd, tr._execute_bwd, tr._execute_kron_fwd, tr._execute_kron_bwd = (
        emulate_trl_tail_fwd, emulate_trl_tail_bwd, emulate_kron_fwd, emulate_kron_bwd)
    saved_cl = cl._execute_fwd, cl._execute_bwd
    cl._execute_fwd, cl._execute_bwd = emulate_cplinear_fwd, emulate_cplinear_bwd
    saved_so = so._execute_spatial_fwd, so._execute_spatial_bwd
    so._execute_spatial_fwd, so._execute_spatial_bwd = emulate_spatial_fwd, emulate_spatial_bwd
    saved_co = co._execute_pack, co._execute_fwd, co._execute_bwd, co._execute_finalize, co._ws_bytes, co._acc_bytes
    co._execute_pack, co._execute_fwd, co._execute_bwd, co._execute_finalize = (
        emulate_cprows_pack, emulate_cprows_fwd, emulate_cprows_bwd, emulate_cprows_finalize)
    co._ws_bytes = co._acc_bytes = lambda d: 64
    saved_k = ko._execute_kr_fwd, ko._execute_kr_bwd
    ko._execute_kr_fwd, ko._execute_kr_bwd = emulate_kr_fwd, emulate_kr_bwd
    saved_r = ro._execute_nchw_to_rows, ro._execute_rows_to_nchw, ro._execute_unfold_rows
    saved_s = ro._execute_skinny_fwd, ro._execute_skinny_bwd, ro._execute_scale_modes, ro._execute_dwconv_rows
    saved_ds = ro._execute_dwsep_taps, ro._execute_dwsep, ro._execute_dwsep_finalize, ro._dwsep_parts
    ro._execute_dwsep_taps, ro._execute_dwsep, ro._execute_dwsep_finalize = emulate_dwsep_taps, emulate_dwsep, emulate_dwsep_finalize
    ro._dwsep_parts = lambda: 2
    saved_ch = ro._execute_chain_prep, ro._execute_chain_post
    ro._execute_chain_prep, ro._execute_chain_post = emulate_chain_prep, emulate_chain_post
    ro._execute_skinny_fwd, ro._execute_skinny_bwd, ro._execute_scale_modes, ro._execute_dwconv_rows = (
        emulate_skinny_fwd, emulate_skinny_bwd, emulate_scale_modes, emulate_dwconv_rows)
    ro._execute_nchw_to_rows, ro._execute_rows_to_nchw, ro._execute_unfold_rows = emulate_nchw_to_rows, emulate_rows_to_nchw, emulate_unfold_rows
    saved_m = mo._execute_mmd
    mo._execute_mmd = emulate_mmd
    saved_t = to._execute_mmd16_fwd, to._execute_mmd16_bwd, to._execute_pack_kmajor
    to._execute_mmd16_fwd, to._execute_mmd16_bwd, to._execute_pack_kmajor = emulate_mmd16_fwd, emulate_mmd16_bwd, emulate_pack_kmajor
    saved_p = po._execute_pointwise, po._execute_pointwise_wgrad
    po._execute_pointwise, po._execute_pointwise_wgrad = emulate_pointwise, emulate_pointwise_wgrad
    saved = ops._execute_contract, ops._execute_conv
    saved_b = bo._execute_btt3_reconstruct, bo._execute_btt3_core_grads, bo._execute_gemm
    ops._execute_contract, ops._execute_conv = emulate_contract, emulate_conv
    bo._execute_btt3_reconstruct, bo._execute_btt3_core_grads, bo._execute_gemm = (
        emulate_btt3_reconstruct, emulate_btt3_core_grads, emulate_gemm)
    del LOG[:]
    try:
        yield LOG
    finally:
        cl._execute_fwd, cl._execute_bwd = saved_cl
        ct._execute_split, ct._execute_kr_bwd_ld, ct._execute_kr_split = saved_ct
        tr._execute_fwd, tr._execute_bwd, tr._execute_kron_fwd, tr._execute_kron_bwd = saved_tr
        so._execute_spatial_fwd, so._execute_spatial_bwd = saved_so
        co._execute_pack, co._execute_fwd, co._execute_bwd, co._execute_finalize, co._ws_bytes, co._acc_bytes = saved_co
        ops._execute_contract, ops._execute_conv = saved
        bo._execute_btt3_reconstruct, bo._execute_btt3_core_grads, bo._execute_gemm = saved_b
        po._execute_pointwise, po._execute_pointwise_wgrad = saved_p
        mo._execute_mmd = saved_m
        ko._execute_kr_fwd, ko._execute_kr_bwd = saved_k
        ro._execute_nchw_to_rows, ro._execute_rows_to_nchw, ro._execute_unfold_rows = saved_r
        ro._execute_skinny_fwd, ro._execute_skinny_bwd, ro._execute_scale_modes, ro._execute_dwconv_rows = saved_s
        ro._execute_dwsep_taps, ro._execute_dwsep, ro._execute_dwsep_finalize, ro._dwsep_parts = saved_ds
        ro._execute_chain_prep, ro._execute_chain_post = saved_ch
        to._execute_mmd16_fwd, to._execute_mmd16_bwd, to._execute_pack_kmajor = saved_t
