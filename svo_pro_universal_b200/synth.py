"""Seeded synthetic EuRoC-shaped inputs for the direct front-end hot path.

Everything here is plain numpy and deterministic in `seed`: the same bytes are fed to the CPU
oracle (tests only) and to the CUDA path. Scene model: a textured plane seen by a pinhole
(optionally radial-tangential) camera; the current image is the reference image inverse-warped
through the plane-induced mapping of a known T_cur_ref, so the alignment optimum is known.

Camera constants: EuRoC cam0 (reference: examples/param/calib/euroc_mono.yaml).
Transformations are 7-vectors (qw qx qy qz tx ty tz), T_a_b maps b-coordinates to a-coordinates.
"""
import numpy as np

EUROC_WIDTH, EUROC_HEIGHT = 752, 480
EUROC_CAM = dict(fx=458.654, fy=457.296, cx=367.215, cy=248.375, k1=0.0, k2=0.0, p1=0.0, p2=0.0,
                 width=EUROC_WIDTH, height=EUROC_HEIGHT, distortion=0)
EUROC_CAM_RADTAN = dict(EUROC_CAM, k1=-0.28340811, k2=0.07395907, p1=0.00019359, p2=1.76187114e-05, distortion=1)
# EuRoC T_B_C (cam0 in body frame); T_cam_imu is its inverse.
_EUROC_T_B_C = np.array([[0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975],
                         [0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768],
                         [-0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949],
                         [0.0, 0.0, 0.0, 1.0]])


# ----------------------------------------------------------------------------------------------
# SE3 helpers (numpy, float64)
def quat_to_rot(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def rot_to_quat(R):
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = np.zeros(4)
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    if q[0] < 0:
        q = -q
    return q / np.linalg.norm(q)


def se3_from_Rt(R, t):
    return np.concatenate([rot_to_quat(R), np.asarray(t, np.float64)])


def se3_to_Rt(T):
    return quat_to_rot(T[:4]), np.asarray(T[4:7], np.float64)


def se3_mul(a, b):
    Ra, ta = se3_to_Rt(a)
    Rb, tb = se3_to_Rt(b)
    return se3_from_Rt(Ra @ Rb, Ra @ tb + ta)


def se3_inv(a):
    R, t = se3_to_Rt(a)
    return se3_from_Rt(R.T, -R.T @ t)


def se3_exp_small(rotvec, t):
    th = np.linalg.norm(rotvec)
    if th < 1e-12:
        R = np.eye(3)
    else:
        k = rotvec / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
    return se3_from_Rt(R, t)


def euroc_T_cam_imu():
    T = np.linalg.inv(_EUROC_T_B_C)
    # re-orthonormalise the 3x3 block (the YAML digits are only ~1e-9 orthonormal)
    u, _, vt = np.linalg.svd(T[:3, :3])
    return se3_from_Rt(u @ vt, T[:3, 3])


IDENTITY = np.array([1.0, 0, 0, 0, 0, 0, 0])


# ----------------------------------------------------------------------------------------------
# camera (vectorised numpy mirror of the pinhole(+radtan) model; used only to *synthesise* inputs)
def cam_project(cam, X):
    x, y = X[..., 0] / X[..., 2], X[..., 1] / X[..., 2]
    if cam["distortion"]:
        k1, k2, p1, p2 = cam["k1"], cam["k2"], cam["p1"], cam["p2"]
        xx, yy, xy = x * x, y * y, x * y
        r2 = xx + yy
        cd = (k1 + k2 * r2) * r2
        x, y = (x + x * cd + p1 * 2 * xy + p2 * (r2 + 2 * xx), y + y * cd + p2 * 2 * xy + p1 * (r2 + 2 * yy))
    return np.stack([cam["fx"] * x + cam["cx"], cam["fy"] * y + cam["cy"]], -1)


def cam_backproject(cam, px):
    x = (px[..., 0] - cam["cx"]) / cam["fx"]
    y = (px[..., 1] - cam["cy"]) / cam["fy"]
    if cam["distortion"]:
        k1, k2, p1, p2 = cam["k1"], cam["k2"], cam["p1"], cam["p2"]
        x0, y0 = x.copy(), y.copy()
        for _ in range(5):
            xx, yy, xy = x * x, y * y, x * y
            r2 = xx + yy
            ic = 1.0 / (1.0 + (k1 + k2 * r2) * r2)
            dx = p1 * 2 * xy + p2 * (r2 + 2 * xx)
            dy = p2 * 2 * xy + p1 * (r2 + 2 * yy)
            x, y = (x0 - dx) * ic, (y0 - dy) * ic
    return np.stack([x, y, np.ones_like(x)], -1)


# ----------------------------------------------------------------------------------------------
# images
def make_image(seed, width=EUROC_WIDTH, height=EUROC_HEIGHT, n_rect=900, blur=0, noise=3):
    """uint8 HxW: random piece-wise constant rectangles (8-64 px) + low-frequency gradient + uniform noise."""
    rng = np.random.default_rng(np.uint64(seed) * np.uint64(2654435761) + np.uint64(12345))
    img = np.full((height, width), 110.0)
    xs = rng.integers(-32, width, n_rect)
    ys = rng.integers(-32, height, n_rect)
    ws = rng.integers(8, 65, n_rect)
    hs = rng.integers(8, 65, n_rect)
    vs = rng.integers(0, 256, n_rect)
    for x, y, w, h, v in zip(xs, ys, ws, hs, vs):
        img[max(y, 0):max(y + h, 0), max(x, 0):max(x + w, 0)] = v
    yy, xx = np.mgrid[0:height, 0:width]
    img = 0.8 * img + 25.0 * np.sin(xx / 97.0 + seed) + 20.0 * np.cos(yy / 61.0 - seed) + 20.0
    for _ in range(blur):
        p = np.pad(img, 1, mode="edge")
        img = (p[:-2, :-2] + p[:-2, 1:-1] + p[:-2, 2:] + p[1:-1, :-2] + p[1:-1, 1:-1] + p[1:-1, 2:]
               + p[2:, :-2] + p[2:, 1:-1] + p[2:, 2:]) / 9.0
    if noise:
        img = img + rng.integers(-noise, noise + 1, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def bilinear(img, u, v):
    """Sample uint8 image at float coords (clamped to the border)."""
    h, w = img.shape
    u = np.clip(u, 0, w - 1.001)
    v = np.clip(v, 0, h - 1.001)
    x0 = np.floor(u).astype(np.int64)
    y0 = np.floor(v).astype(np.int64)
    fx, fy = u - x0, v - y0
    im = img.astype(np.float64)
    return ((1 - fx) * (1 - fy) * im[y0, x0] + fx * (1 - fy) * im[y0, x0 + 1]
            + (1 - fx) * fy * im[y0 + 1, x0] + fx * fy * im[y0 + 1, x0 + 1])


class PlaneScene:
    """Plane n·X = d in the REFERENCE camera frame, textured by the reference image."""

    def __init__(self, ref_img, cam, normal=(0.12, -0.2, 1.0), dist=4.0):
        self.ref_img = ref_img
        self.cam = cam
        n = np.asarray(normal, np.float64)
        self.n = n / np.linalg.norm(n)
        self.d = float(dist)

    def ref_points(self, px):
        """3D points (ref camera frame) seen at ref pixels px[N,2]."""
        f = cam_backproject(self.cam, np.asarray(px, np.float64))
        lam = self.d / (f @ self.n)
        return f * lam[:, None]

    def render(self, T_cur_ref, cam_cur=None):
        """Image seen from the camera at T_cur_ref (inverse warping + bilinear, rounded to uint8)."""
        cam_cur = cam_cur or self.cam
        h, w = cam_cur["height"], cam_cur["width"]
        yy, xx = np.mgrid[0:h, 0:w]
        px = np.stack([xx, yy], -1).astype(np.float64).reshape(-1, 2)
        fc = cam_backproject(cam_cur, px)
        R_rc, t_rc = se3_to_Rt(se3_inv(np.asarray(T_cur_ref, np.float64)))
        d_r = fc @ R_rc.T
        lam = (self.d - self.n @ t_rc) / (d_r @ self.n)
        Xr = d_r * lam[:, None] + t_rc
        ur = cam_project(self.cam, Xr)
        val = bilinear(self.ref_img, ur[:, 0], ur[:, 1])
        return np.clip(np.rint(val), 0, 255).astype(np.uint8).reshape(h, w)


def pick_features(img, n, seed, cell=30, lo=(40, 40), hi=(664, 392), integer=True):
    """Strongest-gradient pixel per `cell`x`cell` bucket inside [lo,hi), then a seeded choice of n."""
    g = img.astype(np.float64)
    gx = np.zeros_like(g)
    gy = np.zeros_like(g)
    gx[:, 1:-1] = g[:, 2:] - g[:, :-2]
    gy[1:-1, :] = g[2:, :] - g[:-2, :]
    # corner-ness: product of |gx| and |gy| smoothed over 3x3 favours corners over straight edges
    m = np.abs(gx) * np.abs(gy) + 0.05 * (gx * gx + gy * gy)
    pts = []
    for y0 in range(lo[1], hi[1], cell):
        for x0 in range(lo[0], hi[0], cell):
            blk = m[y0:min(y0 + cell, hi[1]), x0:min(x0 + cell, hi[0])]
            iy, ix = np.unravel_index(int(np.argmax(blk)), blk.shape)
            if blk[iy, ix] > 0:
                pts.append((x0 + ix, y0 + iy))
    pts = np.array(pts, np.float64)
    rng = np.random.default_rng(seed + 7919)
    if len(pts) > n:
        pts = pts[np.sort(rng.choice(len(pts), n, replace=False))]
    if not integer:
        pts = pts + rng.uniform(-0.4, 0.4, pts.shape)
    return pts


def random_motion(seed, max_rot_deg=2.0, max_trans=0.05):
    rng = np.random.default_rng(seed + 104729)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(0.3, max_rot_deg))
    t = rng.normal(size=3)
    t = t / np.linalg.norm(t) * rng.uniform(0.2, 1.0) * max_trans
    return se3_exp_small(axis * ang, t)


def make_align_pair(seed, n_features=180, cam=None, blur=2, max_rot_deg=2.0, max_trans=0.05, with_imu_extrinsics=True):
    """One synthetic frame pair for SparseImgAlign (BASELINE.json configs[0]).

    Returns dict: ref_img, cur_img (uint8 480x752), cam, T_cam_imu, T_imu_world_ref, T_imu_world_cur_init (= ref pose,
    i.e. identity motion guess), T_cur_ref_gt (camera frame), T_icur_iref_gt, px[N,2], f[N,3] (unit bearing),
    depth[N] (distance from the ref camera centre), eligible[N].
    """
    cam = dict(cam or EUROC_CAM)
    ref_img = make_image(seed, cam["width"], cam["height"], blur=blur)
    scene = PlaneScene(ref_img, cam, normal=(0.12 * np.cos(seed), -0.2 * np.sin(seed * 0.7), 1.0), dist=3.0 + (seed % 5))
    T_cur_ref = random_motion(seed, max_rot_deg, max_trans)
    cur_img = scene.render(T_cur_ref)
    px = pick_features(ref_img, n_features, seed)
    X = scene.ref_points(px)
    depth = np.linalg.norm(X, axis=1)
    f = X / depth[:, None]
    T_cam_imu = euroc_T_cam_imu() if with_imu_extrinsics else IDENTITY.copy()
    rng = np.random.default_rng(seed + 31)
    T_imu_world_ref = se3_exp_small(rng.normal(size=3) * 0.3, rng.normal(size=3))
    T_imu_cam = se3_inv(T_cam_imu)
    T_icur_iref_gt = se3_mul(T_imu_cam, se3_mul(T_cur_ref, T_cam_imu))
    return dict(ref_img=ref_img, cur_img=cur_img, cam=cam, T_cam_imu=T_cam_imu, T_imu_world_ref=T_imu_world_ref,
                T_imu_world_cur_init=T_imu_world_ref.copy(), T_cur_ref_gt=T_cur_ref, T_icur_iref_gt=T_icur_iref_gt,
                px=px, f=f, depth=depth, eligible=np.ones(len(px), np.uint8), scene=scene)


# svo::FeatureType values (reference: src/svo_common/include/svo/common/types.h:60-73)
K_EDGELET_SEED, K_CORNER_SEED, K_MAPPOINT_SEED = 0, 1, 2
K_EDGELET_SEED_CONV, K_CORNER_SEED_CONV, K_MAPPOINT_SEED_CONV = 3, 4, 5
K_EDGELET, K_CORNER, K_MAPPOINT, K_FIXED_LANDMARK, K_OUTLIER = 6, 7, 8, 9, 10


def make_match_set(seed, n_features=2000, cam=None, edgelet_frac=0.25, guess_noise=2.0, max_rot_deg=2.0, max_trans=0.08):
    """Features for findMatchDirect / align2D / align1D (BASELINE.json configs[2]): integer px on textured points,
    level in {0,1,2}, 25 % edgelets with a unit gradient direction, depth from the plane, px guess = truth + U[-n,n]."""
    cam = dict(cam or EUROC_CAM)
    ref_img = make_image(seed, cam["width"], cam["height"], blur=1)
    scene = PlaneScene(ref_img, cam, normal=(0.1, -0.15, 1.0), dist=3.0 + (seed % 4))
    T_cur_ref = random_motion(seed, max_rot_deg, max_trans)
    cur_img = scene.render(T_cur_ref)
    rng = np.random.default_rng(seed + 271)
    px = pick_features(ref_img, n_features, seed, cell=12, lo=(24, 24), hi=(cam["width"] - 24, cam["height"] - 24))
    n = len(px)
    level = rng.integers(0, 3, n).astype(np.int32)
    ftype = np.where(rng.uniform(size=n) < edgelet_frac, K_EDGELET, K_CORNER).astype(np.int32)
    # edgelet gradient direction: local image gradient (normalised), fallback (1,0)
    g = ref_img.astype(np.float64)
    xi, yi = px[:, 0].astype(int), px[:, 1].astype(int)
    gx = g[yi, np.minimum(xi + 1, cam["width"] - 1)] - g[yi, np.maximum(xi - 1, 0)]
    gy = g[np.minimum(yi + 1, cam["height"] - 1), xi] - g[np.maximum(yi - 1, 0), xi]
    nrm = np.hypot(gx, gy)
    grad = np.stack([np.where(nrm > 0, gx / np.maximum(nrm, 1e-12), 1.0), np.where(nrm > 0, gy / np.maximum(nrm, 1e-12), 0.0)], -1)
    X = scene.ref_points(px)
    depth = np.linalg.norm(X, axis=1)
    f = X / depth[:, None]
    R, t = se3_to_Rt(T_cur_ref)
    px_true = cam_project(cam, X @ R.T + t)
    px_guess = px_true + rng.uniform(-guess_noise, guess_noise, px_true.shape)
    return dict(ref_img=ref_img, cur_img=cur_img, cam=cam, T_cur_ref=T_cur_ref, px=px, f=f, depth=depth, level=level,
                type=ftype, grad=grad, px_true=px_true, px_guess=px_guess, scene=scene)


def make_seed_sequence(seed, n_seeds=400, n_obs=8, cam=None, baseline_step=0.01, depth_min=1.5, depth_mean=4.0,
                       edgelet_frac=0.2):
    """One reference keyframe with seeds + n_obs observation frames along a smooth trajectory whose baseline grows
    by `baseline_step` metres per frame (BASELINE.json configs[3]). Seeds are initialised like
    depth_filter_utils::initializeSeeds: mu = 1/depth_mean, sigma2 = (1/depth_min)^2/36, a = b = 10."""
    cam = dict(cam or EUROC_CAM)
    ref_img = make_image(seed, cam["width"], cam["height"], blur=1)
    scene = PlaneScene(ref_img, cam, normal=(0.08, -0.1, 1.0), dist=depth_mean)
    rng = np.random.default_rng(seed + 911)
    direction = np.array([1.0, 0.25, 0.05])
    direction /= np.linalg.norm(direction)
    cur_imgs, T_cur_refs = [], []
    for o in range(n_obs):
        b = baseline_step * (o + 1)
        rot = np.array([0.002, -0.003, 0.001]) * (o + 1)
        T_ref_cur = se3_exp_small(rot, direction * b)
        T_cur_ref = se3_inv(T_ref_cur)
        T_cur_refs.append(T_cur_ref)
        cur_imgs.append(scene.render(T_cur_ref))
    px = pick_features(ref_img, n_seeds, seed, cell=16, lo=(32, 32), hi=(cam["width"] - 32, cam["height"] - 32))
    n = len(px)
    level = rng.integers(0, 3, n).astype(np.int32)
    ftype = np.where(rng.uniform(size=n) < edgelet_frac, K_EDGELET_SEED, K_CORNER_SEED).astype(np.uint8)
    g = ref_img.astype(np.float64)
    xi, yi = px[:, 0].astype(int), px[:, 1].astype(int)
    gx = g[yi, xi + 1] - g[yi, xi - 1]
    gy = g[yi + 1, xi] - g[yi - 1, xi]
    nrm = np.hypot(gx, gy)
    grad = np.stack([np.where(nrm > 0, gx / np.maximum(nrm, 1e-12), 1.0), np.where(nrm > 0, gy / np.maximum(nrm, 1e-12), 0.0)], -1)
    X = scene.ref_points(px)
    depth = np.linalg.norm(X, axis=1)
    f = X / depth[:, None]
    mu_range = 1.0 / depth_min
    state = np.zeros((n, 4))
    state[:, 0] = 1.0 / depth_mean
    state[:, 1] = mu_range * mu_range / 36.0
    state[:, 2:] = 10.0
    return dict(ref_img=ref_img, cur_imgs=cur_imgs, T_cur_ref=np.array(T_cur_refs), cam=cam, px=px, f=f, level=level,
                type=ftype, grad=grad, depth_true=depth, state=state, mu_range=mu_range, scene=scene)


def make_reproject_scene(seed, n_kfs=4, n_per_kf=260, cam=None, integer_scores=False, max_rot_deg=2.5, max_trans=0.12, n_cur=1):
    """A small map for the Reprojector: `n_kfs` keyframes and one current frame looking at the textured plane,
    the keyframes' feature columns (landmarks with observation lists, converged and unconverged seeds, corners and edgelets) as
    the flat tables of svo_reproj_map, and the entries Reprojector::reprojectFrames would visit (each landmark once, every seed).
    World frame = camera frame of the texture image; T_f_w are the frame poses. `n_cur` current frames (cur_imgs, cur_Ts)
    share the map."""
    cam = dict(cam or EUROC_CAM)
    base_img = make_image(seed, cam["width"], cam["height"], blur=1)
    scene = PlaneScene(base_img, cam, normal=(0.1 * np.cos(seed), -0.12, 1.0), dist=3.5 + (seed % 3))
    rng = np.random.default_rng(seed + 4242)
    kf_T = [random_motion(seed * 17 + k, max_rot_deg, max_trans) for k in range(n_kfs)]
    cur_Ts = [random_motion(seed * 17 + 99 + c, max_rot_deg, max_trans) for c in range(n_cur)]
    kf_imgs = [scene.render(T) for T in kf_T]
    cur_imgs = [scene.render(T) for T in cur_Ts]

    def rays_to_world(T_f_w, px):
        """3-D plane points (world) seen at pixels px of the frame with pose T_f_w, and their depth along the ray."""
        f = cam_backproject(cam, px)
        f = f / np.linalg.norm(f, axis=1)[:, None]
        R_wf, t_wf = se3_to_Rt(se3_inv(T_f_w))
        d_w = f @ R_wf.T
        lam = (scene.d - scene.n @ t_wf) / (d_w @ scene.n)
        return d_w * lam[:, None] + t_wf, f, lam

    feats, score, state, point, feat_kf, kf_begin = [], [], [], [], [], [0]
    pt_pos, pt_obs = [], []
    depth_min = 1.5
    mu_range = 1.0 / depth_min
    for k in range(n_kfs):
        img = kf_imgs[k]
        px = pick_features(img, n_per_kf, seed * 31 + k, cell=20, lo=(24, 24), hi=(cam["width"] - 24, cam["height"] - 24))
        Xw, f, lam = rays_to_world(kf_T[k], px)
        g = img.astype(np.float64)
        xi, yi = px[:, 0].astype(int), px[:, 1].astype(int)
        gx = g[yi, xi + 1] - g[yi, xi - 1]
        gy = g[yi + 1, xi] - g[yi - 1, xi]
        nrm = np.hypot(gx, gy)
        grad = np.stack([np.where(nrm > 0, gx / np.maximum(nrm, 1e-12), 1.0), np.where(nrm > 0, gy / np.maximum(nrm, 1e-12), 0.0)], -1)
        for i in range(len(px)):
            u = rng.uniform()
            edgelet = rng.uniform() < 0.2
            level = int(rng.integers(0, 3))
            sc = float(rng.integers(11, 60)) if integer_scores else float(rng.uniform(11.0, 60.0))
            st = np.array([1.0, 1.0, 10.0, 10.0])
            pid = -1
            if u < 0.5:      # landmark
                ftype = K_EDGELET if edgelet else K_CORNER
                pid = len(pt_pos)
                pt_pos.append(Xw[i])
                pt_obs.append([len(feats)])
            elif u < 0.75:   # converged seed: tight around the true inverse depth
                ftype = K_EDGELET_SEED_CONV if edgelet else K_CORNER_SEED_CONV
                st[0] = (1.0 / lam[i]) * (1.0 + rng.normal() * 0.01)
                st[1] = (mu_range / 300.0) ** 2
            else:            # unconverged seed
                ftype = K_EDGELET_SEED if edgelet else K_CORNER_SEED
                if rng.uniform() < 0.3:   # one more good observation converges it (seed::isConverged at thresh 200)
                    st[0] = (1.0 / lam[i]) * (1.0 + rng.normal() * 0.003)
                    st[1] = (mu_range / 199.8) ** 2
                else:
                    st[0] = (1.0 / lam[i]) * (1.0 + rng.normal() * 0.05)
                    st[1] = (0.08 * st[0]) ** 2
            feats.append((px[i], f[i], grad[i], ftype, level))
            score.append(sc); state.append(st); point.append(pid); feat_kf.append(k)
        kf_begin.append(len(feats))
    # a second observation of about half of the landmarks in the next keyframe (not an entry: the reprojector visits a point once)
    n_first = len(feats)
    extra = [[] for _ in range(n_kfs)]
    for pid, X in enumerate(pt_pos):
        if rng.uniform() < 0.5:
            k2 = (feat_kf[pt_obs[pid][0]] + 1) % n_kfs
            R, t = se3_to_Rt(kf_T[k2])
            Xc = R @ X + t
            px2 = cam_project(cam, Xc[None])[0]
            if 30 <= px2[0] < cam["width"] - 30 and 30 <= px2[1] < cam["height"] - 30:
                extra[k2].append((pid, px2, Xc / np.linalg.norm(Xc)))
    # rebuild the tables keyframe by keyframe with the extra observations appended to each keyframe's block
    new_feats, new_score, new_state, new_point, new_kf, new_begin, remap = [], [], [], [], [], [0], {}
    for k in range(n_kfs):
        for i in range(kf_begin[k], kf_begin[k + 1]):
            remap[i] = len(new_feats)
            new_feats.append(feats[i]); new_score.append(score[i]); new_state.append(state[i]); new_point.append(point[i]); new_kf.append(k)
        for pid, px2, f2 in extra[k]:
            pt_obs[pid].append(-len(new_feats) - 1)  # already a new index (negative marks "no remap")
            new_feats.append((px2, f2, np.array([1.0, 0.0]), K_CORNER, int(rng.integers(0, 2))))
            new_score.append(float(rng.uniform(11.0, 60.0))); new_state.append(np.array([1.0, 1.0, 10.0, 10.0]))
            new_point.append(pid); new_kf.append(k)
        new_begin.append(len(new_feats))
    obs_begin, obs_feat = [0], []
    for pid in range(len(pt_pos)):
        for o in pt_obs[pid]:
            obs_feat.append(remap[o] if o >= 0 else -o - 1)
        obs_begin.append(len(obs_feat))
    NF = len(new_feats)
    feat = np.zeros(NF, np.dtype([("px", "<f8", 2), ("f", "<f8", 3), ("grad", "<f8", 2), ("type", "<i4"), ("level", "<i4")]))
    for i, (p, f, gr, t, l) in enumerate(new_feats):
        feat["px"][i], feat["f"][i], feat["grad"][i], feat["type"][i], feat["level"][i] = p, f, gr, t, l
    n_pts = len(pt_pos)
    tables = dict(n_kfs=n_kfs, n_feat=NF, n_points=n_pts, n_obs=len(obs_feat),
                  kf_T_f_w=np.array(kf_T), kf_seed_mu_range=np.full(n_kfs, mu_range), kf_feat_begin=np.array(new_begin, np.int32),
                  feat=feat, feat_score=np.array(new_score), feat_seed_state=np.array(new_state), feat_point=np.array(new_point, np.int32),
                  feat_kf=np.array(new_kf, np.int32), pt_pos=np.array(pt_pos).reshape(n_pts, 3),
                  pt_n_failed=rng.integers(0, 4, n_pts).astype(np.int32), pt_n_succeeded=rng.integers(0, 6, n_pts).astype(np.int32),
                  pt_obs_begin=np.array(obs_begin, np.int32), obs_feat=np.array(obs_feat, np.int32))
    # entries: first observation of every landmark + every seed, keyframe by keyframe (= the original features)
    entry_feat = np.array([remap[i] for i in range(n_first)], np.int32)
    return dict(cam=cam, kf_imgs=kf_imgs, cur_img=cur_imgs[0], cur_T_f_w=cur_Ts[0], cur_imgs=cur_imgs, cur_Ts=np.array(cur_Ts),
                tables=tables, entry_feat=entry_feat, scene=scene)


def make_pose_opt_case(seed, n_per_cam=150, n_cams=1, cam=None, px_noise=0.4, outlier_frac=0.06, edgelet_frac=0.25, rot_err=0.01, trans_err=0.03):
    """A frame bundle for PoseOptimizer::run: n_cams cameras (stereo baseline 11 cm) observing random 3-D points;
    measurements = projections at the true pose + Gaussian pixel noise, a few gross outliers, a few features without a 3-D
    point; the initial IMU pose is the true one perturbed by (rot_err rad, trans_err m)."""
    cam = dict(cam or EUROC_CAM)
    rng = np.random.default_rng(seed + 6000)
    T_cam_imu = [euroc_T_cam_imu()]
    if n_cams == 2:
        T_cam_imu.append(se3_mul(se3_exp_small(np.zeros(3), np.array([-0.11, 0.0, 0.0])), T_cam_imu[0]))
    T_imu_world_true = se3_exp_small(rng.normal(size=3) * 0.2, rng.normal(size=3))
    px, f, grad, level, ftype, xyz, has, fcam = [], [], [], [], [], [], [], []
    for c in range(n_cams):
        T_f_w = se3_mul(T_cam_imu[c], T_imu_world_true)
        R, t = se3_to_Rt(T_f_w)
        p = np.stack([rng.uniform(20, cam["width"] - 20, n_per_cam), rng.uniform(20, cam["height"] - 20, n_per_cam)], 1)
        d = rng.uniform(2.0, 8.0, n_per_cam)
        fb = cam_backproject(cam, p)
        fb /= np.linalg.norm(fb, axis=1)[:, None]
        Xc = fb * d[:, None]
        Xw = (Xc - t) @ R          # R^T (Xc - t)
        lv = rng.integers(0, 3, n_per_cam)
        meas = p + rng.normal(size=p.shape) * px_noise * (1 << lv)[:, None]
        out = rng.uniform(size=n_per_cam) < outlier_frac
        meas[out] += rng.uniform(-25, 25, (int(out.sum()), 2))
        fm = cam_backproject(cam, meas)
        fm /= np.linalg.norm(fm, axis=1)[:, None]
        g = rng.normal(size=(n_per_cam, 2)); g /= np.linalg.norm(g, axis=1)[:, None]
        ty = np.where(rng.uniform(size=n_per_cam) < edgelet_frac, K_EDGELET, K_CORNER)
        hx = rng.uniform(size=n_per_cam) > 0.05
        ty = np.where(hx, ty, K_OUTLIER)
        px.append(meas); f.append(fm); grad.append(g); level.append(lv); ftype.append(ty); xyz.append(Xw); has.append(hx); fcam.append(np.full(n_per_cam, c))
    T_init = se3_mul(se3_exp_small(rng.normal(size=3) * rot_err, rng.normal(size=3) * trans_err), T_imu_world_true)
    cat = np.concatenate
    return dict(cam=cam, T_cam_imu=T_cam_imu, T_imu_world_true=T_imu_world_true, T_imu_world_init=T_init, px=cat(px), f=cat(f), grad=cat(grad),
                level=cat(level).astype(np.int32), type=cat(ftype).astype(np.int32), xyz_world=cat(xyz), has_xyz=cat(has).astype(np.uint8),
                feat_cam=cat(fcam).astype(np.int32))
