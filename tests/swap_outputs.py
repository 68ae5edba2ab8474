"""Helper of tests/test_gpu_ref_swap.py, run in its OWN process (a C++ exception inside the swap library would otherwise take the whole
pytest session down): every golden-producing entry point of oracle/ref_frontend_wrapper.cpp evaluated through
oracle/_ref/libfrontend_swap.so, written to an .npz.   python tests/swap_outputs.py out.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from oracle import orc  # noqa: E402


def main(path):
    orc.lib()
    orc.use_frontend_lib("swap")
    L = orc.ref_frontend_lib()
    assert L is not None, "oracle/_ref/libfrontend_swap.so is missing"
    out = {}
    for k, v in helpers.frontend_outputs(orc, "ref").items():   # the "ref" entry points, now resolved inside libfrontend_swap.so
        out["fe_" + k] = np.asarray(v)
    for k, v in helpers.reproject_frames_reference(orc).items():
        out["rp_" + k] = np.asarray(v)
    if hasattr(L, "ref_stereo_triangulation_compute"):
        for k, v in helpers.stereo_tri_reference(orc).items():
            out["st_" + k] = np.asarray(v)
    # the scalar leaves: filter updates and tau against the oracle's values on the same inputs
    rng = np.random.default_rng(2)
    rows = []
    for _ in range(300):
        st = np.array([rng.uniform(0.05, 1), rng.uniform(1e-6, 0.05), rng.uniform(1, 30), rng.uniform(1, 30)])
        z, tau2 = rng.uniform(0.01, 1.5), 10.0 ** rng.uniform(-8, -1)
        a, b, c, d = st.copy(), st.copy(), st.copy(), st.copy()
        ra = orc.lib().orc_update_filter_vogiatzis(z, tau2, 0.66, orc._f64(a)); rb = L.ref_update_filter_vogiatzis(z, tau2, 0.66, orc._f64(b))
        rc = orc.lib().orc_update_filter_gaussian(z, tau2, orc._f64(c)); rd = L.ref_update_filter_gaussian(z, tau2, orc._f64(d))
        rows.append(np.concatenate([[ra, rb, rc, rd], a, b, c, d]))
    out["leaf_filter"] = np.array(rows)
    taus = []
    for _ in range(100):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        T = np.concatenate([q, rng.normal(size=3) * 0.3])
        f = rng.normal(size=3); f /= np.linalg.norm(f)
        z = rng.uniform(0.3, 12)
        taus.append([orc.lib().orc_compute_tau(orc._f64(T), orc._f64(f), z, 0.00218), L.ref_compute_tau(orc._f64(T), orc._f64(f), z, 0.00218)])
    out["leaf_tau"] = np.array(taus)
    np.savez(path, **out)


if __name__ == "__main__":
    main(sys.argv[1])
