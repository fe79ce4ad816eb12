"""Regenerates tests/golden/*.npz from the reference's in-tree datasets (run HERE, where /root/reference exists).

Each fixture holds (a) the parsed input graph in '.g2o payload' form and (b) what the oracle - the CPU
restatement linked against the reference's vendored CSparse - produces for it: block-AMD permutation, nnz(L),
per-iteration chi2 / lambda, final estimates.  The GPU parity tests read these files; they never read
/root/reference.   usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle_binding import GN, LM, Oracle, fnv1a64  # noqa: E402

DATA = "/root/reference/data"
CASES = [  # name, file, algorithm, iterations
    ("manhattan3500", "2d/manhattan3500/manhattanOlson3500.g2o", GN, 10),
    ("intel", "2d/intel/intel.g2o", GN, 10),
    ("sphere_bignoise", "3d/sphere/sphere_bignoise_vertex3.g2o", LM, 10),
    ("garage", "3d/garage/parking-garage.g2o", LM, 10),
]
VT = {"VERTEX_SE2": 0, "VERTEX_SE3:QUAT": 1, "VERTEX_CAM": 2, "VERTEX_XYZ": 3}
ET = {"EDGE_SE2": 0, "EDGE_SE3:QUAT": 1, "EDGE_PROJECT_P2MC": 2}


def parse(path):
    """file-order list of ('v'|'e'|'f', ...) records"""
    v_ids, v_kind, v_pay, e_a, e_b, e_kind, e_pay, fixed, order = [], [], [], [], [], [], [], [], []
    for line in open(path):
        t = line.split()
        if not t or t[0].startswith("#"):
            continue
        if t[0] == "FIX":
            fixed += [int(x) for x in t[1:]]
        elif t[0] in VT:
            v_ids.append(int(t[1])); v_kind.append(VT[t[0]]); v_pay.append([float(x) for x in t[2:]]); order.append(0)
        elif t[0] in ET:
            e_a.append(int(t[1])); e_b.append(int(t[2])); e_kind.append(ET[t[0]]); e_pay.append([float(x) for x in t[3:]]); order.append(1)
    return dict(v_ids=np.array(v_ids, np.int32), v_kind=np.array(v_kind, np.int32), v_pay=np.array(v_pay),
                e_a=np.array(e_a, np.int32), e_b=np.array(e_b, np.int32), e_kind=np.array(e_kind, np.int32),
                e_pay=np.array(e_pay), fixed=np.array(fixed, np.int32), order=np.array(order, np.int8))


def main():
    for name, rel, algo, iters in CASES:
        path = os.path.join(DATA, rel)
        inp = parse(path)
        o = Oracle()
        assert o.load(path)
        gauge = o.setup_cli(True)
        assert o.initialize_optimization()
        # first linear system, before anything moves
        o.algorithm_init(); o.build_structure()
        chi0 = o.compute_active_errors()
        o.build_system()
        b0 = o.b()
        n, st = o.optimize(algo, iters)
        ids, kinds, hidx, flags = o.vertices()
        est = np.stack([np.pad(o.vertex_estimate(i), (0, 12))[:12] for i in ids])
        perm = o.block_perm()
        out = dict(inp)
        out.update(gauge=gauge, algorithm=algo, iterations=iters, done=n, chi2_initial=chi0, b_initial=b0,
                   chi2=np.array([s.chi2 for s in st]), lam=np.array([s.lambda_ for s in st]),
                   lev_iters=np.array([s.levenberg_iterations for s in st]), final_ids=ids, final_kinds=kinds,
                   final_hidx=hidx, final_flags=flags, final_est=est, perm=perm, perm_hash=fnv1a64(perm),
                   lnz=o.lnz(), dims=np.array(list(o.dims().values())))
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, "gauge", gauge, "iters", n, "chi2", st[n - 1].chi2, "lnz", o.lnz(), fnv1a64(perm))


if __name__ == "__main__":
    main()
