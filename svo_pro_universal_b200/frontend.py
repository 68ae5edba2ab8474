"""The full direct front-end of one tracked stereo frame, batched over B independent stereo frame pairs (BASELINE.json
configs[4]) — every stage a C-ABI call on device-resident arrays, no host round trip between them:

    pyramid of the two new frames                     svo_cuda_pyr_build            (frame_handler_base.cpp:184-186)
 -> SparseImgAlign::run on the 2-camera bundle        svo_cuda_sparse_align         (frame_handler_base.cpp:610-643)
 -> Reprojector::reprojectFrames per camera           svo_cuda_reproject_match      (frame_handler_base.cpp:646-743)
 -> PoseOptimizer::run on the matched features        svo_cuda_pose_optimize        (frame_handler_base.cpp:746-790)
 -> DepthFilter::updateSeeds of the last keyframe     svo_cuda_update_seeds         (frame_handler_stereo.cpp:82)
 -> FastGradDetector on the new left frame            svo_cuda_fastgrad_detect      (depth_filter.cpp:255-365, new keyframe; the
                                                                                      default detector, svo_factory.cpp:292-295)

The per-pair units are independent, so a batch shards over GPUs by contiguous blocks of pairs (shard.partition) with no
collective in the path. PyTorch only holds the device arrays and reshapes the aligner's poses into the reprojector's input.
"""
import numpy as np

from . import capi, synth

W, H, N_LEVELS = synth.EUROC_WIDTH, synth.EUROC_HEIGHT, 5
T_C1_C0 = synth.se3_exp_small(np.zeros(3), np.array([-0.11, 0.0, 0.0]))  # EuRoC-like 11 cm stereo baseline


def make_stereo_scene(seed, n0=180, n1=150):
    """One synthetic stereo frame pair around synth.make_align_pair(seed): images, features and extrinsics of both cameras,
    the frames' ground-truth poses, and seeds (unconverged, depth-filter initial state) on the left reference frame."""
    d = synth.make_align_pair(seed, n_features=n0)
    scene, cam = d["scene"], d["cam"]
    ref1 = scene.render(T_C1_C0)
    cur1 = scene.render(synth.se3_mul(T_C1_C0, d["T_cur_ref_gt"]))
    px1 = synth.pick_features(ref1, n1, seed + 5)
    f1 = synth.cam_backproject(cam, px1)
    R01, t01 = synth.se3_to_Rt(synth.se3_inv(T_C1_C0))
    lam = (scene.d - scene.n @ t01) / ((f1 @ R01.T) @ scene.n)
    X1 = f1 * lam[:, None]
    depth1 = np.linalg.norm(X1, axis=1)
    T_cam_imu = [d["T_cam_imu"], synth.se3_mul(T_C1_C0, d["T_cam_imu"])]
    T_f_w_ref = [synth.se3_mul(T, d["T_imu_world_ref"]) for T in T_cam_imu]
    # seeds on the left reference frame: a second set of corners, inverse depth 5 % off, sigma of a fresh seed
    rng = np.random.default_rng(seed + 77)
    pxs = synth.pick_features(d["ref_img"], 120, seed + 9, cell=40)
    Xs = scene.ref_points(pxs)
    ds = np.linalg.norm(Xs, axis=1)
    mu_range = 1.0 / 1.5
    st = np.zeros((len(pxs), 4))
    st[:, 0] = (1.0 / ds) * (1.0 + rng.normal(size=len(pxs)) * 0.05)
    st[:, 1] = (0.08 * st[:, 0]) ** 2
    st[:, 2:] = 10.0
    return dict(cam=cam, imgs=dict(r0=d["ref_img"], r1=ref1, c0=d["cur_img"], c1=cur1), px=[d["px"], px1],
                f=[d["f"], X1 / depth1[:, None]], depth=[d["depth"], depth1], T_cam_imu=T_cam_imu, T_f_w_ref=T_f_w_ref,
                T_imu_world_ref=d["T_imu_world_ref"], T_cur_ref_gt=d["T_cur_ref_gt"], seed_px=pxs, seed_f=Xs / ds[:, None],
                seed_state=st, seed_mu_range=mu_range, level_rng=seed)


class StereoFrontendBatch:
    """B stereo frame pairs made of `scenes` (unique synthetic stereo scenes, tiled) resident on one GPU."""

    def __init__(self, ctx, scenes, B, device, max_features=180, stereo_triangulation=False):
        import torch
        self.torch, self.ctx, self.B, self.dev, self.F = torch, ctx, B, device, max_features
        self.with_stereo = bool(stereo_triangulation)
        # ONE stream for the library's kernels and for torch's glue ops (slicing the aligner's poses, resetting the grid):
        # the context is switched to a torch stream and every step runs with that stream current
        self.stream = torch.cuda.Stream(device=device)
        ctx.set_stream(self.stream.cuda_stream)
        U = len(scenes)
        self.scenes, self.sid = scenes, np.arange(B) % U
        sid = self.sid
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
        self.cam = capi.Camera.from_dict(scenes[0]["cam"])
        # pyramids: ref and cur batches hold the left frames in [0, B) and the right frames in [B, 2B)
        self.ref = capi.Pyramid(ctx, 2 * B, W, H, N_LEVELS)
        self.cur = capi.Pyramid(ctx, 2 * B, W, H, N_LEVELS)
        self.ref.upload(t(np.stack([scenes[s]["imgs"]["r0"] for s in sid] + [scenes[s]["imgs"]["r1"] for s in sid])))
        self.ref.build()
        self.cur.upload(t(np.stack([scenes[s]["imgs"]["c0"] for s in sid] + [scenes[s]["imgs"]["c1"] for s in sid])))
        self.frame_idx = t(np.stack([np.arange(B), B + np.arange(B)], 1).astype(np.int32))  # [B][cam]
        # ---- sparse alignment inputs
        F = max_features
        upx = np.zeros((U, 2, F, 2)); uf = np.zeros((U, 2, F, 3)); udep = np.ones((U, 2, F)); uel = np.zeros((U, 2, F), np.uint8)
        unf = np.zeros((U, 2), np.int32)
        for s_, sc in enumerate(scenes):
            for c in range(2):
                n = len(sc["px"][c])
                upx[s_, c, :n], uf[s_, c, :n], udep[s_, c, :n], uel[s_, c, :n], unf[s_, c] = sc["px"][c], sc["f"][c], sc["depth"][c], 1, n
        px, f, dep, el, nf = upx[sid], uf[sid], udep[sid], uel[sid], unf[sid]
        self.align_in = {k: t(v) for k, v in dict(px=px, f=f, depth=dep, eligible=el, n_features=nf).items()}
        Tw = np.stack([scenes[s]["T_imu_world_ref"] for s in sid])
        self.T_imu_world_ref, self.T_imu_world_cur = t(Tw), t(Tw.copy())
        self.T_cam_imu = np.stack(scenes[0]["T_cam_imu"])
        self.align_opt = capi.sparse_align_options(estimate_illumination_gain=1, estimate_illumination_offset=1)  # vio_stereo.yaml
        self.d_align = torch.zeros(B * capi.ALIGN_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=device)
        # ---- reprojector tables: keyframe 2i+c = ref frame of camera c of pair i; the left features are landmarks (one
        # observation each), the right features converged seeds
        blocks = []  # per unique scene: the feature rows of its two keyframes, left camera first
        for sc in scenes:
            n0, n1 = len(sc["px"][0]), len(sc["px"][1])
            ft = np.concatenate([capi.make_features(sc["px"][c], sc["f"][c], np.tile([1.0, 0.0], (len(sc["px"][c]), 1)),
                                                    np.full(len(sc["px"][c]), synth.K_CORNER if c == 0 else synth.K_CORNER_SEED_CONV, np.int32),
                                                    np.zeros(len(sc["px"][c]), np.int32)) for c in range(2)])
            st = np.tile([1.0, 1e-6, 10.0, 10.0], (n0 + n1, 1))
            st[:, 0] = 1.0 / np.concatenate(sc["depth"])
            R, tt = synth.se3_to_Rt(synth.se3_inv(sc["T_f_w_ref"][0]))
            blocks.append(dict(n0=n0, n1=n1, feat=ft, score=np.concatenate([np.linspace(60.0, 11.0, n0), np.linspace(60.0, 11.0, n1)]),
                               state=st, Xw=(sc["f"][0] * sc["depth"][0][:, None]) @ R.T + tt))
        n0s, n1s = np.array([blocks[s_]["n0"] for s_ in sid]), np.array([blocks[s_]["n1"] for s_ in sid])
        fbase = np.concatenate([[0], np.cumsum(n0s + n1s)])[:-1]          # first feature of pair i
        pbase = np.concatenate([[0], np.cumsum(n0s)])[:-1]                # first landmark of pair i
        cat = lambda k: np.concatenate([blocks[s_][k] for s_ in sid])
        feat = cat("feat")
        # local index of every feature inside its pair, and which camera it belongs to
        pair_of_feat = np.repeat(np.arange(B), n0s + n1s)
        local = np.arange(len(feat)) - fbase[pair_of_feat]
        is_left = local < n0s[pair_of_feat]
        feat_kf = (2 * pair_of_feat + (~is_left)).astype(np.int32)
        point = np.where(is_left, pbase[pair_of_feat] + local, -1).astype(np.int32)
        n_pts = int(n0s.sum())
        obs_feat = np.nonzero(is_left)[0].astype(np.int32)               # landmark p is observed by the p-th left feature
        kf_T = np.stack([scenes[s_]["T_f_w_ref"][c] for s_ in sid for c in range(2)])
        kf_idx = np.array([c * B + i for i in range(B) for c in range(2)], np.int32)
        # entries of current frame j = 2i+c: the features of keyframe 2i+c, in order
        cnt = np.stack([n0s, n1s], 1).reshape(-1)
        eb = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        ef = np.arange(len(feat), dtype=np.int32)                        # features are already stored keyframe by keyframe
        self.tables = dict(n_kfs=2 * B, n_feat=len(feat), n_points=n_pts, n_obs=n_pts, kf_T_f_w=t(kf_T),
                           kf_seed_mu_range=t(np.full(2 * B, 1.0 / 1.5)), kf_frame_idx=t(kf_idx),
                           feat=t(feat.view(np.uint8)), feat_score=t(cat("score")), feat_seed_state=t(cat("state")),
                           feat_point=t(point), feat_kf=t(feat_kf), pt_pos=t(cat("Xw")),
                           pt_n_failed=t(np.zeros(n_pts, np.int32)), pt_n_succeeded=t(np.zeros(n_pts, np.int32)),
                           pt_obs_begin=t(np.arange(n_pts + 1, dtype=np.int32)), obs_feat=t(obs_feat))
        self.host_tables = dict(feat=feat, feat_kf=feat_kf, feat_point=point, eb=eb)
        # ---- pose optimiser: bundle i = the entries of both cameras of pair i; an entry is a measurement once it is matched.
        # Its 3-D point: the landmark, or the converged seed's position from its reference keyframe (pose_optimizer.cpp:123-137)
        R1 = [synth.se3_to_Rt(synth.se3_inv(sc["T_f_w_ref"][1])) for sc in scenes]
        xyz_blocks = [np.concatenate([blocks[u]["Xw"], (scenes[u]["f"][1] * scenes[u]["depth"][1][:, None]) @ R1[u][0].T + R1[u][1]]) for u in range(U)]
        self.po_xyz = t(np.concatenate([xyz_blocks[s_] for s_ in sid]))
        self.po_cam = t((~is_left).astype(np.int32))
        self.po_begin = t(eb[::2].copy())
        self.po_type = t(feat["type"].astype(np.int32))
        self.po_ftrs = torch.zeros((len(feat), 8), dtype=torch.float64, device=device)   # svo_feature rows (64 bytes)
        self.po_opt = capi.pose_optimizer_options()
        self.d_po = torch.zeros(B * capi.POSE_OPT_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=device)
        self.d_po_outlier = torch.zeros(len(feat), dtype=torch.uint8, device=device)
        q_ic, t_ic = synth.se3_inv(scenes[0]["T_cam_imu"][0])[:4], synth.se3_inv(scenes[0]["T_cam_imu"][0])[4:]
        self.q_ic, self.t_ic = t(q_ic), t(t_ic)
        self.entry_begin, self.entry_feat = t(eb), t(ef)
        self.n_entries = int(eb[-1])
        self.reproj_cur_idx = t(np.array([c * B + i for i in range(B) for c in range(2)], np.int32))  # frame j = 2i+c
        self.n_in = torch.zeros(2 * B, dtype=torch.int32, device=device)
        self.n_cells = capi.grid_cells(W, H, 30)[0]
        self.occ0 = torch.zeros((2 * B, self.n_cells), dtype=torch.uint8, device=device)
        self.occ = self.occ0.clone()
        self.reproj_opt = capi.reprojector_options(max_n_features=120)
        self.d_reproj = torch.zeros(self.n_entries * capi.REPROJ_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=device)
        self.d_rstats = torch.zeros(2 * B * capi.REPROJ_STATS_DTYPE.itemsize, dtype=torch.uint8, device=device)
        # ---- depth filter: the seeds of the left reference frames observed by the new left frames (pose of the last
        # optimisation = ground truth here, as the reference updates seeds with the previous frame's final pose)
        rng = np.random.default_rng(1)
        sp = [scenes[s] for s in sid]
        ns = [len(q["seed_px"]) for q in sp]
        sft = capi.make_features(np.concatenate([q["seed_px"] for q in sp]), np.concatenate([q["seed_f"] for q in sp]),
                                 np.tile([1.0, 0.0], (sum(ns), 1)), np.full(sum(ns), synth.K_CORNER_SEED, np.int32),
                                 rng.integers(0, 3, sum(ns)).astype(np.int32))
        self.S = len(sft)
        self.seed_ftrs = t(sft.view(np.uint8))
        self.seed_types0, self.seed_state0 = t(np.full(self.S, synth.K_CORNER_SEED, np.uint8)), t(np.concatenate([q["seed_state"] for q in sp]))
        self.seed_types, self.seed_state = self.seed_types0.clone(), self.seed_state0.clone()
        self.seed_mu_range = t(np.full(self.S, 1.0 / 1.5))
        pair_of_seed = np.repeat(np.arange(B), ns).astype(np.int32)
        self.seed_ref_idx = t(pair_of_seed)                        # left ref frame of the pair
        self.seed_obs_frame = t(pair_of_seed.reshape(1, -1).copy())  # left cur frame of the pair
        self.seed_obs_T = t(sid[pair_of_seed].astype(np.int32).reshape(1, -1))
        self.seed_T = t(np.stack([q["T_cur_ref_gt"] for q in scenes]))
        self.mopt, self.dopt = capi.matcher_options(), capi.depth_filter_options()
        # ---- detector
        self.det_opt = capi.detector_options()
        self.threshold_secondary = 100
        self.d_corners = torch.zeros(B * self.n_cells * capi.CORNER_DTYPE.itemsize, dtype=torch.uint8, device=device)
        self.d_edgelets = torch.zeros_like(self.d_corners)
        # ---- optional keyframe stage: StereoTriangulation::compute on the features just detected in the new left frame
        # (stereo_triangulation.cpp:87-137; the visiting order is corners then edgelets in cell order instead of the reference's
        # std::random_shuffle, entries of empty cells are holes) against the new right frame
        if self.with_stereo:
            nE = 2 * self.n_cells
            self.st_T_f1f0 = synth.se3_mul(self.T_cam_imu[1], synth.se3_inv(self.T_cam_imu[0]))
            self.st_begin = t((np.arange(B + 1) * nE).astype(np.int32))
            self.st_want = t(np.full(B, 120, np.int32))
            self.st_slot0 = t(np.zeros(B, np.int32))
            self.st_f0 = t(np.arange(B, dtype=np.int32))
            self.st_f1 = t((B + np.arange(B)).astype(np.int32))
            self.st_ftrs = torch.zeros((B * nE, 8), dtype=torch.float64, device=device)
            self.st_mopt = capi.matcher_options(max_epi_search_steps=500, subpix_refinement=1)
            self.d_stereo = torch.zeros(B * nE * capi.STEREO_RESULT_DTYPE.itemsize, dtype=torch.uint8, device=device)
            self.d_stereo_stats = torch.zeros(B * capi.STEREO_STATS_DTYPE.itemsize, dtype=torch.uint8, device=device)
            self.st_Twc = torch.zeros((B, 7), dtype=torch.float64, device=device)
        torch.cuda.synchronize(device)   # the uploads above ran on torch's default stream

    def release(self):
        """Give the context its own stream back (the batch object must not be stepped afterwards)."""
        self.torch.cuda.synchronize(self.dev)
        self.ctx.set_stream(0)

    STAGES = ("pyramid", "sparse_align", "reproject", "pose_optimize", "update_seeds", "fastgrad_detect")
    STEREO_STAGE = "stereo_triangulate"  # stage 6 when the batch was built with stereo_triangulation=True

    def _imu_pose(self, T_f_w0):
        """T_imu_world = T_imu_cam0 * T_f_w of the left camera ([B,7] quaternion + translation), on the device."""
        torch = self.torch
        a, b = self.q_ic, T_f_w0[:, :4]
        aw, ax, ay, az = a[0], a[1], a[2], a[3]
        bw, bx, by, bz = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
        q = torch.stack([aw * bw - ax * bx - ay * by - az * bz, aw * bx + ax * bw + ay * bz - az * by,
                         aw * by - ax * bz + ay * bw + az * bx, aw * bz + ax * by - ay * bx + az * bw], 1)
        v = T_f_w0[:, 4:7]
        u = a[1:4].expand_as(v)
        uv = torch.cross(u, v, dim=1)
        rot = v + 2.0 * (aw * uv + torch.cross(u, uv, dim=1))
        return torch.cat([q, rot + self.t_ic], 1).contiguous()

    def step(self, mark=None):
        """One pass of the whole front-end over the batch; everything stays on the device. mark(i) is called after stage i
        (bench: records a CUDA event on the stream)."""
        with self.torch.cuda.stream(self.stream):
            self._step(mark or (lambda i: None))

    def _step(self, mark):
        torch = self.torch
        B, a = self.B, self.align_in
        self.cur.build()
        mark(0)
        capi.sparse_align(self.ctx, [self.ref, self.ref], [self.cur, self.cur], [self.cam, self.cam], self.T_cam_imu, self.T_imu_world_ref,
                          self.T_imu_world_cur, a["n_features"], a["px"], a["f"], a["depth"], a["eligible"], self.align_opt,
                          ref_frame_idx=self.frame_idx, cur_frame_idx=self.frame_idx, results=self.d_align)
        # T_f_w of the new frames: doubles 7..21 of every svo_align_result -> [2B][7], frame j = 2*pair + camera
        stride = capi.ALIGN_RESULT_DTYPE.itemsize // 8
        cur_T = self.d_align.view(torch.float64).view(B, stride)[:, 7:21].reshape(2 * B, 7).contiguous()
        mark(1)
        self.occ.copy_(self.occ0)
        capi.reproject_match(self.ctx, self.ref, self.cur, self.cam, self.cam, self.tables, cur_T, self.n_in, self.entry_begin, self.entry_feat,
                             self.occ, self.reproj_opt, cur_frame_idx=self.reproj_cur_idx, results=self.d_reproj, stats=self.d_rstats)
        mark(2)
        # the matched entries become the measurements of the pose optimiser: (px, f, grad) of the match, the candidate's type
        nE = self.n_entries
        r64 = self.d_reproj.view(torch.float64).view(nE, 17)
        r32 = self.d_reproj.view(torch.int32).view(nE, 34)
        self.po_ftrs[:, 0:7] = r64[:, 2:9]
        f32 = self.po_ftrs.view(torch.int32).view(nE, 16)
        f32[:, 14] = self.po_type
        f32[:, 15] = r32[:, 29]
        self.po_has = (r32[:, 26] == capi.REPROJ_MATCHED).to(torch.uint8)
        self.po_T0 = self._imu_pose(cur_T.view(B, 2, 7)[:, 0])
        capi.pose_optimize(self.ctx, [self.cam, self.cam], self.T_cam_imu, self.po_T0, self.po_begin, self.po_ftrs.view(torch.uint8).view(-1),
                           self.po_cam, self.po_xyz, self.po_has, self.po_opt, results=self.d_po, outlier=self.d_po_outlier)
        mark(3)
        self.seed_types.copy_(self.seed_types0); self.seed_state.copy_(self.seed_state0)
        self.n_seed_ok, _ = capi.update_seeds(self.ctx, self.ref, self.cur, self.cam, self.cam, self.seed_ftrs, self.seed_types, self.seed_state,
                                              self.seed_mu_range, self.seed_obs_frame, self.seed_obs_T, self.seed_T, self.mopt, self.dopt,
                                              ref_frame_idx=self.seed_ref_idx, want_match_results=False)
        mark(4)
        capi.fastgrad_detect(self.ctx, self.cur, self.det_opt, self.threshold_secondary, first=0, count=B, corners_out=self.d_corners,
                             edgelets_out=self.d_edgelets)
        mark(5)
        if self.with_stereo:
            self._stereo_stage(cur_T)
            mark(6)

    def _stereo_stage(self, cur_T):
        """Device-side glue: the per-cell corners / edgelets of the new left frames become svo_feature records (holes where a cell is
        empty), the aligner's left pose is inverted, then ONE svo_cuda_stereo_triangulate call for the whole batch."""
        torch = self.torch
        B, nc = self.B, self.n_cells
        cam = self.scenes[0]["cam"]
        ci = torch.cat([self.d_corners.view(torch.int32).view(B, nc, 5), self.d_edgelets.view(torch.int32).view(B, nc, 5)], 1)
        cf = torch.cat([self.d_corners.view(torch.float32).view(B, nc, 5), self.d_edgelets.view(torch.float32).view(B, nc, 5)], 1)
        thr = torch.cat([torch.full((nc,), float(self.det_opt.threshold), device=self.dev),
                         torch.full((nc,), float(self.threshold_secondary), device=self.dev)]).to(torch.float32)
        ftype = torch.cat([torch.full((nc,), 7, device=self.dev), torch.full((nc,), 6, device=self.dev)]).to(torch.int32)  # kCorner, kEdgelet
        valid = cf[..., 3] > thr[None, :]
        x, y = ci[..., 0].to(torch.float64), ci[..., 1].to(torch.float64)
        bx, by = (x - cam["cx"]) / cam["fx"], (y - cam["cy"]) / cam["fy"]  # pinhole without distortion (the chain's cameras)
        n = torch.sqrt(bx * bx + by * by + 1.0)
        rec = self.st_ftrs.view(B, 2 * nc, 8)
        rec[..., 0], rec[..., 1] = x, y
        rec[..., 2], rec[..., 3], rec[..., 4] = bx / n, by / n, 1.0 / n
        rec[..., 5], rec[..., 6] = torch.cos(cf[..., 4]).to(torch.float64), torch.sin(cf[..., 4]).to(torch.float64)
        r32 = self.st_ftrs.view(torch.int32).view(B, 2 * nc, 16)
        r32[..., 14] = torch.where(valid, ftype[None, :].expand(B, -1), torch.full_like(ci[..., 0], -1))
        r32[..., 15] = ci[..., 2]
        # T_world_cam of the new left frames = inverse of the aligner's T_f_w
        T = cur_T.view(B, 2, 7)[:, 0]
        qi = T[:, :4] * torch.tensor([1.0, -1.0, -1.0, -1.0], dtype=torch.float64, device=self.dev)
        u, v, aw = qi[:, 1:4], T[:, 4:7], qi[:, 0:1]
        uv = torch.cross(u, v, dim=1)
        self.st_Twc[:, :4] = qi
        self.st_Twc[:, 4:] = -(v + 2.0 * (aw * uv + torch.cross(u, uv, dim=1)))
        capi.stereo_triangulate(self.ctx, self.cur, self.cur, self.cam, self.cam, self.st_T_f1f0, self.st_Twc, self.st_begin,
                                self.st_ftrs.view(torch.uint8).view(-1), self.st_want, self.st_slot0, self.st_mopt, frame0_idx=self.st_f0,
                                frame1_idx=self.st_f1, results=self.d_stereo, stats=self.d_stereo_stats)

    def results(self):
        """Host copies of every stage's outputs (numpy)."""
        self.ctx.synchronize()
        self.torch.cuda.synchronize(self.dev)
        return dict(align=self.d_align.cpu().numpy().view(capi.ALIGN_RESULT_DTYPE),
                    reproj=self.d_reproj.cpu().numpy().view(capi.REPROJ_RESULT_DTYPE),
                    reproj_stats=self.d_rstats.cpu().numpy().view(capi.REPROJ_STATS_DTYPE), occupancy=self.occ.cpu().numpy(),
                    seed_types=self.seed_types.cpu().numpy(), seed_state=self.seed_state.cpu().numpy(),
                    n_seed_ok=int(self.n_seed_ok.item()), pose_opt=self.d_po.cpu().numpy().view(capi.POSE_OPT_RESULT_DTYPE),
                    pose_opt_outlier=self.d_po_outlier.cpu().numpy(), pose_opt_T0=self.po_T0.cpu().numpy(),
                    pose_opt_has=self.po_has.cpu().numpy(), pose_opt_xyz=self.po_xyz.cpu().numpy(),
                    corners=self.d_corners.cpu().numpy().view(capi.CORNER_DTYPE).reshape(self.B, self.n_cells),
                    edgelets=self.d_edgelets.cpu().numpy().view(capi.CORNER_DTYPE).reshape(self.B, self.n_cells),
                    entry_begin=self.entry_begin.cpu().numpy(),
                    **(dict(stereo=self.d_stereo.cpu().numpy().view(capi.STEREO_RESULT_DTYPE).reshape(self.B, 2 * self.n_cells),
                            stereo_stats=self.d_stereo_stats.cpu().numpy().view(capi.STEREO_STATS_DTYPE),
                            stereo_ftrs=self.st_ftrs.cpu().numpy().view(capi.FEATURE_DTYPE).reshape(self.B, 2 * self.n_cells),
                            stereo_Twc=self.st_Twc.cpu().numpy(), align_T_f_w=self.d_align.cpu().numpy().view(capi.ALIGN_RESULT_DTYPE)["T_f_w"])
                       if self.with_stereo else {}))
