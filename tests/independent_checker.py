"""Independent sequential checker of the raster adapter (numpy fp64).  TEST INFRASTRUCTURE ONLY.

A second, separately written statement of what GaussianRasterizer.__call__ computes
[REF tinysplat/splatting/rasterize.py:26-62]: it shares NO code with oracle/ and none with the
kernels, has no tiles-of-pixels batching, no autograd and no sorting by composite keys — every
Gaussian is visited once in (depth, id) order and composited onto the pixels of the tile rectangle
its 3-sigma square touches.  Gradients are obtained by central finite differences of this forward
(tests/test_independent_checker.py), so the kernels' analytic backward, the oracle's autograd
backward and this file's numerical derivative are three separate routes to the same numbers.

Stated algorithm (the published 3DGS/gsplat-legacy constants): +0.3 on the
projected covariance diagonal, radius = ceil(3 sqrt(lambda_max)) with the discriminant floored at
0.1, frustum clamp 1.3 in the EWA Jacobian, w + 1e-6, pixel centres at +0.5, near plane 0.01,
alpha = min(0.999, o exp(-sigma)), contributions with alpha < 1/255 skipped, a pixel stops before
the Gaussian that would bring its transmittance to <= 1e-4."""
import numpy as np

Y0 = 0.28209479177387814
Y1 = 0.4886025119029199


def _sh_colour(deg, d, coef):
    """coef [K,3]; d unit vector.  Real spherical harmonics, bands 0..deg."""
    x, y, z = d
    b = [Y0]
    if deg >= 1:
        b += [-Y1 * y, Y1 * z, -Y1 * x]
    if deg >= 2:
        b += [1.0925484305920792 * x * y, -1.0925484305920792 * y * z,
              0.31539156525252005 * (2 * z * z - x * x - y * y),
              -1.0925484305920792 * x * z, 0.5462742152960396 * (x * x - y * y)]
    if deg >= 3:
        b += [-0.5900435899266435 * y * (3 * x * x - y * y), 2.890611442640554 * x * y * z,
              -0.4570457994644658 * y * (4 * z * z - x * x - y * y),
              0.3731763325901154 * z * (2 * z * z - 3 * x * x - 3 * y * y),
              -0.4570457994644658 * x * (4 * z * z - x * x - y * y),
              1.445305721320277 * z * (x * x - y * y), -0.5900435899266435 * x * (x * x - 3 * y * y)]
    if deg >= 4:
        raise NotImplementedError("checker covers SH degree <= 3")
    return np.asarray(b) @ coef[:len(b)]


def render(p, view, proj, fx, fy, W, H, deg, window=None, dir_means=None):
    """p: dict of float64 numpy arrays means[N,3], scales[N,3] (log), quats[N,4] (w,x,y,z),
    opacities[N,1] (logit), colors_dc[N,3], colors_rest[N,K-1,3], background[3].
    window = (x0, y0, x1, y1) in PIXELS (half-open) or None.  Returns rgb[h,w,3] (clamped to <= 1),
    depth[h,w] (composited over background[0]) for the window.  dir_means: the means the SH view
    directions are taken from (default p["means"]); the stated backward gives view directions NO
    gradient, so a finite difference w.r.t. the means holds them fixed."""
    view = np.asarray(view, np.float64)
    full = np.asarray(proj, np.float64) @ view
    x0, y0, x1, y1 = window if window is not None else (0, 0, W, H)
    tbx, tby = -(-W // 16), -(-H // 16)
    N = p["means"].shape[0]
    Rv, tv = view[:3, :3], view[:3, 3]
    items = []
    for i in range(N):
        mu = p["means"][i]
        t = Rv @ mu + tv
        if not t[2] > 0.01:
            continue
        w_, qx, qy, qz = p["quats"][i] / np.linalg.norm(p["quats"][i])
        R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - w_ * qz), 2 * (qx * qz + w_ * qy)],
                      [2 * (qx * qy + w_ * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - w_ * qx)],
                      [2 * (qx * qz - w_ * qy), 2 * (qy * qz + w_ * qx), 1 - 2 * (qx * qx + qy * qy)]])
        S = np.diag(np.exp(p["scales"][i]) ** 2)
        Sigma = R @ S @ R.T
        lx, ly = 1.3 * 0.5 * W / fx, 1.3 * 0.5 * H / fy
        cx_ = t[2] * min(lx, max(-lx, t[0] / t[2]))
        cy_ = t[2] * min(ly, max(-ly, t[1] / t[2]))
        J = np.array([[fx / t[2], 0.0, -fx * cx_ / t[2] ** 2], [0.0, fy / t[2], -fy * cy_ / t[2] ** 2]])
        A = J @ Rv
        cov = A @ Sigma @ A.T
        a, b, c = cov[0, 0] + 0.3, cov[0, 1], cov[1, 1] + 0.3
        det = a * c - b * b
        if det == 0 or not np.isfinite(det):
            continue
        mid = 0.5 * (a + c)
        rad = np.ceil(3.0 * np.sqrt(mid + np.sqrt(max(0.1, mid * mid - det))))
        ph = full @ np.append(mu, 1.0)
        wq = ph[3] + 1e-6
        if wq == 0 or not rad > 0:
            continue
        px = 0.5 * W * ph[0] / wq + W / 2 - 0.5
        py = 0.5 * H * ph[1] / wq + H / 2 - 0.5
        lo_x = min(max(0, int(np.floor((px - rad) / 16))), tbx)
        hi_x = min(max(0, int(np.floor((px + rad) / 16 + 1))), tbx)
        lo_y = min(max(0, int(np.floor((py - rad) / 16))), tby)
        hi_y = min(max(0, int(np.floor((py + rad) / 16 + 1))), tby)
        if hi_x <= lo_x or hi_y <= lo_y:
            continue
        d = (mu if dir_means is None else dir_means[i]) - tv   # the reference's view direction [REF rasterize.py:77-79]
        coef = np.concatenate([p["colors_dc"][i][None, :], p["colors_rest"][i]], axis=0)
        rgb = np.maximum(_sh_colour(deg, d / np.linalg.norm(d), coef) + 0.5, 0.0)
        op = 1.0 / (1.0 + np.exp(-p["opacities"][i, 0]))
        items.append((t[2], i, px, py, c / det, -b / det, a / det, op, rgb,
                      (16 * lo_x, 16 * lo_y, min(16 * hi_x, W), min(16 * hi_y, H))))
    items.sort(key=lambda it: (it[0], it[1]))                 # front to back, ties by Gaussian id
    h, w = y1 - y0, x1 - x0
    T = np.ones((h, w))
    acc = np.zeros((h, w, 4))
    done = np.zeros((h, w), bool)
    jj, ii = np.meshgrid(np.arange(x0, x1) + 0.5, np.arange(y0, y1) + 0.5)
    for dep, i, px, py, ca, cb, cc, op, rgb, (rx0, ry0, rx1, ry1) in items:
        ax0, ay0, ax1, ay1 = max(rx0, x0) - x0, max(ry0, y0) - y0, min(rx1, x1) - x0, min(ry1, y1) - y0
        if ax1 <= ax0 or ay1 <= ay0:
            continue
        sl = (slice(ay0, ay1), slice(ax0, ax1))
        dx, dy = px - jj[sl], py - ii[sl]
        sig = 0.5 * (ca * dx * dx + cc * dy * dy) + cb * dx * dy
        alpha = np.minimum(0.999, op * np.exp(-sig))
        hit = (sig >= 0) & (alpha >= 1.0 / 255.0) & ~done[sl]
        nT = T[sl] * (1.0 - alpha)
        stop = hit & (nT <= 1e-4)
        done[sl] |= stop
        upd = hit & ~stop
        col = np.append(rgb, dep)
        acc[sl] += np.where(upd, alpha * T[sl], 0.0)[..., None] * col
        T[sl] = np.where(upd, nT, T[sl])
    bg = np.asarray(p["background"], np.float64)
    rgb_img = np.minimum(acc[..., :3] + T[..., None] * bg, 1.0)
    depth_img = acc[..., 3] + T * bg[0]
    return rgb_img, depth_img
