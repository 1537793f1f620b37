"""Closed-form (no autograd) statement of the fused warp-SSIM kernels' arithmetic, in torch.

This is the derivation the CUDA kernels in fsnet_b200/csrc/warp_ssim.cu implement line by line:
explicit bilinear-gather coordinate derivatives, SSIM partial derivatives with respect to the 3x3 box
sums, the adjoint of the reflect-padded box filter, the transposed bilinear up-sample and the pose
Jacobian.  tests/test_loss_math.py checks it against autograd through the oracle, so an error in the
derivation is caught on the CPU before it is baked into a kernel.  Test infrastructure only.
"""
import torch
import torch.nn.functional as F

C1 = 0.01 ** 2
C2 = 0.03 ** 2


def upsample_weights(n_in, n_out):
    """align_corners=True bilinear: per output index -> (i0, i1, lambda1)."""
    if n_out > 1:
        scale = torch.tensor((n_in - 1) / (n_out - 1), dtype=torch.float32)
    else:
        scale = torch.tensor(0.0)
    src = scale * torch.arange(n_out, dtype=torch.float32)
    i0 = src.floor().long().clamp(max=n_in - 1)
    i1 = torch.where(i0 < n_in - 1, i0 + 1, i0)
    lam = src - i0.float()
    return i0, i1, lam


def upsample(depth_s, H, W):
    y0, y1, ly = upsample_weights(depth_s.shape[-2], H)
    x0, x1, lx = upsample_weights(depth_s.shape[-1], W)
    d = depth_s[:, 0]
    ly = ly.view(1, -1, 1)
    lx = lx.view(1, 1, -1)
    top = (1 - lx) * d[:, y0][:, :, x0] + lx * d[:, y0][:, :, x1]
    bot = (1 - lx) * d[:, y1][:, :, x0] + lx * d[:, y1][:, :, x1]
    return (1 - ly) * top + ly * bot


def upsample_transpose(gD, hs, ws):
    B, H, W = gD.shape
    y0, y1, ly = upsample_weights(hs, H)
    x0, x1, lx = upsample_weights(ws, W)
    out = torch.zeros(B, hs * ws, dtype=gD.dtype)
    ly = ly.view(1, -1, 1)
    lx = lx.view(1, 1, -1)
    for yi, wy in ((y0, 1 - ly), (y1, ly)):
        for xi, wx in ((x0, 1 - lx), (x1, lx)):
            idx = (yi.view(-1, 1) * ws + xi.view(1, -1)).view(1, -1).expand(B, -1)
            out.scatter_add_(1, idx, (gD * wy * wx).reshape(B, -1))
    return out.view(B, 1, hs, ws)


def box3_reflect(x):
    """3x3 SUM with reflect-101 padding."""
    return F.avg_pool2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), 3, 1) * 9.0


def box3_reflect_adjoint(w):
    """Adjoint of box3_reflect: zero-padded 3x3 sum plus the doubled border rows/columns."""
    def along(w, dim):
        n = w.shape[dim]
        z = F.pad(w, (1, 1, 0, 0) if dim == -1 else (0, 0, 1, 1))
        s = z.narrow(dim, 0, n) + z.narrow(dim, 1, n) + z.narrow(dim, 2, n)
        extra = torch.zeros_like(w)
        extra.narrow(dim, 1, 1).add_(w.narrow(dim, 0, 1))
        extra.narrow(dim, n - 2, 1).add_(w.narrow(dim, n - 1, 1))
        return s + extra
    return along(along(w, -1), -2)


def camera(P2, T):
    """invK (3x3) and P = (K T)[:3] in fp64, cast to fp32 -- what fsnet_camera_setup produces."""
    B = P2.shape[0]
    K = torch.zeros(B, 4, 4, dtype=torch.float64)
    K[:, :3, :3] = P2[:, :3, :3].double()
    K[:, 3, 3] = 1
    invK = torch.linalg.inv(K)[:, :3, :3].float()
    P = torch.matmul(K.float(), T.float())[:, :3, :]
    return invK, P, K.float()


def warp_terms(D, invK, P, src, mask_f32, H, W):
    """Per pixel: pred [B,3,H,W], valid [B,H,W], and the derivative bundle."""
    B = D.shape[0]
    ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    k = invK.view(B, 3, 3, 1, 1)
    r = k[:, :, 0] * xs + k[:, :, 1] * ys + k[:, :, 2]             # [B,3,H,W]
    cam = r * D.unsqueeze(1)
    Pm = P.view(B, 3, 4, 1, 1)
    p = Pm[:, :, 0] * cam[:, 0:1] + Pm[:, :, 1] * cam[:, 1:2] + Pm[:, :, 2] * cam[:, 2:3] + Pm[:, :, 3]
    a = Pm[:, :, 0] * r[:, 0:1] + Pm[:, :, 1] * r[:, 1:2] + Pm[:, :, 2] * r[:, 2:3]      # dp/dD
    zden = p[:, 2] + 1e-7
    u, v = p[:, 0] / zden, p[:, 1] / zden
    ix_raw = ((u / (W - 1) - 0.5) * 2 + 1) / 2 * (W - 1)
    iy_raw = ((v / (H - 1) - 0.5) * 2 + 1) / 2 * (H - 1)
    mx = ((ix_raw > 0) & (ix_raw < W - 1)).float()
    my = ((iy_raw > 0) & (iy_raw < H - 1)).float()
    ix = ix_raw.clamp(0, W - 1)
    iy = iy_raw.clamp(0, H - 1)
    x0, y0 = ix.floor(), iy.floor()
    fx, fy = ix - x0, iy - y0
    x0l, y0l = x0.long(), y0.long()
    x1l, y1l = x0l + 1, y0l + 1
    inx1 = (x1l <= W - 1).float()
    iny1 = (y1l <= H - 1).float()
    x1c, y1c = x1l.clamp(max=W - 1), y1l.clamp(max=H - 1)
    flat = src.reshape(B, 3, H * W)

    def g(yl, xl):
        return torch.gather(flat, 2, (yl * W + xl).view(B, 1, -1).expand(-1, 3, -1)).view(B, 3, H, W)
    nw, ne = g(y0l, x0l), g(y0l, x1c) * inx1.unsqueeze(1)
    sw, se = g(y1c, x0l) * iny1.unsqueeze(1), g(y1c, x1c) * (inx1 * iny1).unsqueeze(1)
    fx_, fy_ = fx.unsqueeze(1), fy.unsqueeze(1)
    pred = nw * (1 - fx_) * (1 - fy_) + ne * fx_ * (1 - fy_) + sw * (1 - fx_) * fy_ + se * fx_ * fy_
    dpred_dix = (ne - nw) * (1 - fy_) + (se - sw) * fy_
    dpred_diy = (sw - nw) * (1 - fx_) + (se - ne) * fx_
    du_dD = (a[:, 0] * zden - p[:, 0] * a[:, 2]) / (zden * zden)
    dv_dD = (a[:, 1] * zden - p[:, 1] * a[:, 2]) / (zden * zden)
    # nearest, zeros padding, validity of the overlap mask
    xn, yn = torch.round(ix_raw), torch.round(iy_raw)      # torch.round = half-to-even = nearbyint
    inb = (xn >= 0) & (xn <= W - 1) & (yn >= 0) & (yn <= H - 1)
    mval = torch.gather(mask_f32.reshape(B, -1), 1, (yn.clamp(0, H - 1).long() * W + xn.clamp(0, W - 1).long()).view(B, -1)).view(B, H, W)
    valid = inb & (mval == 1)
    h = torch.cat([cam, torch.ones(B, 1, H, W)], 1)
    return dict(pred=pred, valid=valid, dix=dpred_dix * mx.unsqueeze(1), diy=dpred_diy * my.unsqueeze(1),
                du_dD=du_dD, dv_dD=dv_dD, zden=zden, p=p, h=h)


def ssim_and_partials(x, y):
    """SSIM loss map of (x=pred, y=target) and d ssim / d (Sx, Sxx, Sxy) at every pixel."""
    Sx, Sy = box3_reflect(x), box3_reflect(y)
    Sxx, Syy, Sxy = box3_reflect(x * x), box3_reflect(y * y), box3_reflect(x * y)
    mx_, my_ = Sx / 9, Sy / 9
    sx, sy, sxy = Sxx / 9 - mx_ * mx_, Syy / 9 - my_ * my_, Sxy / 9 - mx_ * my_
    A1, A2 = 2 * mx_ * my_ + C1, 2 * sxy + C2
    B1, B2 = mx_ * mx_ + my_ * my_ + C1, sx + sy + C2
    Q = A1 * A2 / (B1 * B2)
    raw = (1 - Q) / 2
    out = raw.clamp(0, 1)
    live = ((raw >= 0) & (raw <= 1)).float()
    dQ_dmx = ((2 * my_ * A2 - 2 * my_ * A1) * B1 * B2 - A1 * A2 * (2 * mx_ * B2 - 2 * mx_ * B1)) / (B1 * B2) ** 2
    dS_x = -0.5 * live * dQ_dmx / 9
    dS_xx = -0.5 * live * (-(A1 * A2) / (B1 * B2 * B2)) / 9
    dS_xy = -0.5 * live * (2 * A1 / (B1 * B2)) / 9
    return out, dS_x, dS_xx, dS_xy


def photometric(pred, target):
    ss, a, b, c = ssim_and_partials(pred, target)
    l1 = (target - pred).abs().mean(1)
    return 0.85 * ss.mean(1) + 0.15 * l1, (a, b, c)


def scale_forward_backward(depth_s, target, srcs, cams, mask, ident, overlapped, grad_scale_total, motion_mask=None):
    """One scale of the reprojection loss and its explicit gradients.

    depth_s [B,1,hs,ws]; srcs = [src(+1), src(-1)]; cams = [(invK, P), ...]; mask [B,H,W] or None;
    ident [B,2,H,W] (identity photometric terms + noise) or None for the motion-mask branch.
    Returns loss_s (photometric part), grad depth_s, grad P per frame.
    """
    B, _, H, W = target.shape
    hs, ws = depth_s.shape[-2:]
    D = upsample(depth_s, H, W)
    m32 = torch.ones(B, H, W) if mask is None else mask.float()
    terms, photos, partials = [], [], []
    for f in range(2):
        invK, P = cams[f]
        t = warp_terms(D, invK, P, srcs[f], m32, H, W)
        ph, part = photometric(t["pred"], target)
        if overlapped:
            ph = torch.where(t["valid"], ph, torch.full_like(ph, 100.0))
        terms.append(t)
        photos.append(ph)
        partials.append(part)
    reproj = torch.stack(photos, 1)
    if ident is not None:
        comb = torch.cat([ident, reproj], 1)
        val, idx = comb.min(1)
        win = [(idx == 2), (idx == 3)]
        gate = torch.ones(B, H, W)
    else:
        val, idx = reproj.min(1)
        win = [(idx == 0), (idx == 1)]
        gate = 1 - motion_mask
    maskv = torch.ones(B, H, W, dtype=torch.float64) if mask is None else mask.double()
    den = maskv.sum() + 1e-6
    loss = (val.double() * maskv).sum() / den
    g_pix = (grad_scale_total * maskv / den).float() * gate        # d total / d val(p)
    gD = torch.zeros(B, H, W)
    gP = []
    for f in range(2):
        t = terms[f]
        a, b, c = partials[f]
        gf = g_pix * win[f].float()
        if overlapped:
            gf = gf * t["valid"].float()
        w = (0.85 / 3) * gf.unsqueeze(1)
        Ga, Gb, Gc = box3_reflect_adjoint(w * a), box3_reflect_adjoint(w * b), box3_reflect_adjoint(w * c)
        gpred = Ga + 2 * t["pred"] * Gb + target * Gc
        gpred = gpred + (0.15 / 3) * gf.unsqueeze(1) * (-torch.sign(target - t["pred"]))
        gu = (gpred * t["dix"]).sum(1)
        gv = (gpred * t["diy"]).sum(1)
        gD = gD + gu * t["du_dD"] + gv * t["dv_dD"]
        z = t["zden"]
        row0 = (gu / z).unsqueeze(1) * t["h"]
        row1 = (gv / z).unsqueeze(1) * t["h"]
        row2 = (-(gu * t["p"][:, 0] + gv * t["p"][:, 1]) / (z * z)).unsqueeze(1) * t["h"]
        gP.append(torch.stack([row0.sum((2, 3)), row1.sum((2, 3)), row2.sum((2, 3))], 1))     # [B,3,4]
    return loss, upsample_transpose(gD, hs, ws), gP, idx


def smooth_forward_backward(disp, color, weight):
    """Edge-aware smoothness on mean-normalised disparity and d/d disp (explicit)."""
    B, _, h, w = disp.shape
    mean = disp.mean((2, 3), keepdim=True)
    nd = disp / (mean + 1e-7)
    wx = torch.exp(-(color[:, :, :, :-1] - color[:, :, :, 1:]).abs().mean(1, True))
    wy = torch.exp(-(color[:, :, :-1, :] - color[:, :, 1:, :]).abs().mean(1, True))
    dx = nd[:, :, :, :-1] - nd[:, :, :, 1:]
    dy = nd[:, :, :-1, :] - nd[:, :, 1:, :]
    nx, ny = B * h * (w - 1), B * (h - 1) * w
    loss = ((dx.abs() * wx).sum() / nx + (dy.abs() * wy).sum() / ny) * weight
    gx = torch.sign(dx) * wx * (weight / nx)
    gy = torch.sign(dy) * wy * (weight / ny)
    gnd = torch.zeros_like(nd)
    gnd[:, :, :, :-1] += gx
    gnd[:, :, :, 1:] -= gx
    gnd[:, :, :-1, :] += gy
    gnd[:, :, 1:, :] -= gy
    m = mean + 1e-7
    gdisp = gnd / m - (gnd * disp).sum((2, 3), keepdim=True) / (m * m) / (h * w)
    return loss, gdisp
