"""Packing of synthetic frame pairs into the flat arrays the C ABI takes (host side, numpy)."""
import numpy as np


def pack_align_batch(pairs, max_features=None):
    """pairs: list of dicts from synth.make_align_pair (mono). Returns dict of numpy arrays shaped for
    capi.sparse_align with n_cams = 1: px [B,1,F,2], f [B,1,F,3], depth [B,1,F], eligible [B,1,F], n_features [B,1],
    T_imu_world_ref/cur [B,7], ref_imgs/cur_imgs [B,H,W] uint8."""
    B = len(pairs)
    F = max_features or max(len(p["px"]) for p in pairs)
    out = dict(px=np.zeros((B, 1, F, 2)), f=np.zeros((B, 1, F, 3)), depth=np.ones((B, 1, F)),
               eligible=np.zeros((B, 1, F), np.uint8), n_features=np.zeros((B, 1), np.int32),
               T_imu_world_ref=np.zeros((B, 7)), T_imu_world_cur=np.zeros((B, 7)))
    h, w = pairs[0]["ref_img"].shape
    out["ref_imgs"] = np.zeros((B, h, w), np.uint8)
    out["cur_imgs"] = np.zeros((B, h, w), np.uint8)
    for i, p in enumerate(pairs):
        n = min(len(p["px"]), F)
        out["px"][i, 0, :n] = p["px"][:n]
        out["f"][i, 0, :n] = p["f"][:n]
        out["depth"][i, 0, :n] = p["depth"][:n]
        out["eligible"][i, 0, :n] = p["eligible"][:n]
        out["n_features"][i, 0] = n
        out["T_imu_world_ref"][i] = p["T_imu_world_ref"]
        out["T_imu_world_cur"][i] = p["T_imu_world_cur_init"]
        out["ref_imgs"][i] = p["ref_img"]
        out["cur_imgs"][i] = p["cur_img"]
    out["T_cam_imu"] = np.asarray(pairs[0]["T_cam_imu"], np.float64).reshape(1, 7)
    out["cam"] = pairs[0]["cam"]
    return out


def tile_batch(packed, B):
    """Repeat a packed batch of K unique pairs up to B pairs (bench workloads: K unique synthetic pairs, tiled)."""
    K = packed["px"].shape[0]
    idx = np.arange(B) % K
    out = {}
    for k, v in packed.items():
        if isinstance(v, np.ndarray) and v.shape[:1] == (K,) and k not in ("T_cam_imu",):
            out[k] = np.ascontiguousarray(v[idx])
        else:
            out[k] = v
    return out
