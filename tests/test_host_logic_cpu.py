"""CPU tests (-m "not gpu"): host-side logic — synthetic generator determinism, batch packing, sharding over 2 gloo ranks."""
import os
import sys

import numpy as np

from svo_pro_universal_b200 import batch, shard, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_well_formed():
    a, b = synth.make_align_pair(5), synth.make_align_pair(5)
    assert np.array_equal(a["ref_img"], b["ref_img"]) and np.array_equal(a["cur_img"], b["cur_img"])
    assert a["ref_img"].shape == (480, 752) and a["ref_img"].dtype == np.uint8
    assert len(a["px"]) == 180
    assert (a["px"][:, 0] >= 40).all() and (a["px"][:, 0] < 664).all() and (a["px"][:, 1] >= 40).all() and (a["px"][:, 1] < 392).all()
    np.testing.assert_allclose(np.linalg.norm(a["f"], axis=1), 1.0, atol=1e-12)
    assert (a["depth"] > 1.0).all() and (a["depth"] < 12.0).all()
    # T_icur_iref_gt is T_cur_ref conjugated by the camera-IMU extrinsics
    T = synth.se3_mul(a["T_cam_imu"], synth.se3_mul(a["T_icur_iref_gt"], synth.se3_inv(a["T_cam_imu"])))
    np.testing.assert_allclose(T, a["T_cur_ref_gt"], atol=1e-12)


def test_plane_scene_is_geometrically_consistent():
    d = synth.make_align_pair(9)
    X = d["scene"].ref_points(d["px"])
    np.testing.assert_allclose(synth.cam_project(d["cam"], X), d["px"], atol=1e-9)
    R, t = synth.se3_to_Rt(d["T_cur_ref_gt"])
    pc = synth.cam_project(d["cam"], X @ R.T + t)
    inside = (pc[:, 0] > 8) & (pc[:, 0] < 744) & (pc[:, 1] > 8) & (pc[:, 1] < 472)
    a = synth.bilinear(d["ref_img"], d["px"][inside, 0], d["px"][inside, 1])
    b = synth.bilinear(d["cur_img"], pc[inside, 0], pc[inside, 1])
    assert np.median(np.abs(a - b)) < 6.0  # same surface point, same intensity up to interpolation / quantisation


def test_radtan_backproject_inverts_project():
    cam = synth.EUROC_CAM_RADTAN
    # 5 fixed-point iterations (radial_tangential_distortion.h:80-95) converge well near the centre, less so far out
    px = np.array([[250.0, 180.0], [367.0, 248.0], [480.0, 330.0]])
    np.testing.assert_allclose(synth.cam_project(cam, synth.cam_backproject(cam, px)), px, atol=2e-2)


def test_pack_and_tile_batch():
    pairs = [synth.make_align_pair(s, n_features=50 + 10 * s) for s in range(3)]
    pk = batch.pack_align_batch(pairs)
    assert pk["px"].shape == (3, 1, 70, 2) and list(pk["n_features"][:, 0]) == [50, 60, 70]
    assert pk["eligible"][0, 0, 50:].sum() == 0
    tiled = batch.tile_batch(pk, 8)
    assert tiled["px"].shape[0] == 8 and np.array_equal(tiled["px"][3], pk["px"][0]) and tiled["T_cam_imu"].shape == (1, 7)


def test_partition_covers_all_units():
    for n in (0, 1, 7, 4096, 8191):
        for w in (1, 2, 3, 8):
            blocks = [shard.partition(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from svo_pro_universal_b200 import shard as sh
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    lo, hi = sh.partition(n, world, rank)
    dt = np.dtype([("T", "<f8", 7), ("n", "<i4"), ("pad", "<i4")])
    local = np.zeros(hi - lo, dt)
    local["n"] = np.arange(lo, hi)
    local["T"][:, 0] = np.arange(lo, hi) * 0.5
    full = sh.gather_to_rank0(local, n)
    if rank == 0:
        q.put((full["n"].tolist(), full["T"][:, 0].tolist()))
    else:
        assert full is None
    dist.destroy_process_group()


def test_two_rank_gloo_shard_and_gather():
    """N>1 path on CPU: two gloo ranks each own a contiguous block; only the final gather is collective."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ns, ts = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert ns == list(range(11)) and ts == [0.5 * i for i in range(11)]
