"""Seeded synthetic inputs of the shapes BASELINE.json's configs name (no datasets ship, no network).

Each generator returns a dict of numpy arrays in ".g2o payload" form (the numbers that follow the ids on the
text line), so the same arrays feed the product (SparseOptimizer.add_vertices/add_edges), the oracle, and
`write_g2o`.  Input generation is host-side test/bench plumbing, not part of the hot path.

  sphere(nodes_per_level, laps, ...)   examples/sphere/create_sphere.cpp:95-184 (reference generator)
  venice_like(cams, points, ...)       SURVEY.md section 8d config 3/4: ring of inward-looking cameras, windowed visibility
  expmap_ba(cams, points, ...)         the same scene as VERTEX_SE3:EXPMAP / EDGE_PROJECT_XYZ2UV:EXPMAP
  ba_demo(pixel_noise, outliers, ...)  the scene of the reference's examples/ba/ba_demo.cpp
  landmark_slam_2d(poses, landmarks)   SE2 odometry + EdgeSE2PointXY sightings (the set-up of examples/tutorial_slam2d)
  landmark_slam_3d(poses, landmarks)   SE3 odometry + EdgeSE3PointXYZ sightings through a ParameterSE3Offset
"""
import numpy as np

# vertex / edge kinds of the C-ABI (include/g2o_b200.h).  Literal copies, so that this module has no import besides numpy:
# bench.py's reference arm loads it by file path without importing the package (and with it libg2o_b200.so)
VERTEX_SE2, VERTEX_SE3, VERTEX_CAM, VERTEX_XYZ, VERTEX_SE3_EXPMAP, VERTEX_XY = 0, 1, 2, 3, 4, 5
EDGE_SE2, EDGE_SE3, EDGE_P2MC, EDGE_XYZ2UV, EDGE_SE2_XY, EDGE_SE3_XYZ = 0, 1, 2, 3, 4, 5


# ---------------------------------------------------------------- quaternion helpers (x y z w), vectorised
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by,
                     aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], axis=-1)


def _qconj(q):
    return q * np.array([-1.0, -1.0, -1.0, 1.0])


def _qrot(q, v):
    qv = np.concatenate([v, np.zeros(v.shape[:-1] + (1,))], axis=-1)
    return _qmul(_qmul(q, qv), _qconj(q))[..., :3]


def _compose(qa, ta, qb, tb):
    return _qmul(qa, qb), ta + _qrot(qa, tb)


def _prefix_compose(q, t):
    """inclusive scan of SE3 composition (Hillis-Steele): out[i] = T0*T1*...*Ti"""
    q, t = q.copy(), t.copy()
    n, step = len(q), 1
    while step < n:
        q2, t2 = _compose(q[:-step], t[:-step], q[step:], t[step:])
        q[step:], t[step:] = q2, t2
        step *= 2
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    return q, t


def sphere(nodes_per_level=50, laps=50, radius=100.0, sigma_t=0.01, sigma_r=0.005, seed=2500):
    """SE3 pose graph on a sphere (create_sphere.cpp:95-184). Defaults = sphere2500 (2500 poses, 9799 edges)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = nodes_per_level * laps
    ids = np.arange(n)
    nn = ids % nodes_per_level
    az = -np.pi + 2 * nn * np.pi / nodes_per_level
    ay = -0.5 * np.pi + (ids + 1) * np.pi / n       # `id` is already incremented in the reference (:102-105)
    qz = np.stack([0 * az, 0 * az, np.sin(az / 2), np.cos(az / 2)], axis=-1)
    qy = np.stack([0 * ay, np.sin(ay / 2), 0 * ay, np.cos(ay / 2)], axis=-1)
    q = _qmul(qz, qy)
    t = _qrot(q, np.tile(np.array([radius, 0.0, 0.0]), (n, 1)))
    # edges: odometry chain, then loop closures to the previous lap in the reference's order
    # (for f: for nn: for dn in -1,0,+1, skipping +1 on the last lap; :131-147)
    v0_parts, v1_parts = [ids[:-1]], [ids[1:]]
    for f in range(1, laps):
        dns = np.array([-1, 0] if f == laps - 1 else [-1, 0, 1])
        base = np.arange(nodes_per_level)
        v0_parts.append(np.repeat((f - 1) * nodes_per_level + base, len(dns)))
        v1_parts.append(((f * nodes_per_level + base)[:, None] + dns[None, :]).reshape(-1))
    v0 = np.concatenate(v0_parts)
    v1 = np.concatenate(v1_parts)
    # ground-truth relative transforms, then noise (:155-175)
    qrel = _qmul(_qconj(q[v0]), q[v1])
    trel = _qrot(_qconj(q[v0]), t[v1] - t[v0])
    E = len(v0)
    v = rng.normal(0.0, sigma_r, (E, 3))
    qw = np.maximum(0.0, 1.0 - np.linalg.norm(v, axis=1))
    qn = np.concatenate([v, qw[:, None]], axis=1)
    qn /= np.linalg.norm(qn, axis=1, keepdims=True)
    qmeas = _qmul(qrel, qn)
    tmeas = trel + rng.normal(0.0, sigma_t, (E, 3))
    info = np.zeros((6, 6))
    info[:3, :3] = np.eye(3) / sigma_t ** 2
    info[3:, 3:] = np.eye(3) / sigma_r ** 2
    iu = np.array([info[i, j] for i in range(6) for j in range(i, 6)])
    edge_payload = np.concatenate([tmeas, qmeas, np.tile(iu, (E, 1))], axis=1)
    # initial guess: concatenate the (noisy) odometry (:178-184)
    qo = np.concatenate([q[:1], qmeas[:n - 1]], axis=0)
    to = np.concatenate([t[:1], tmeas[:n - 1]], axis=0)
    qi, ti = _prefix_compose(qo, to)
    vert_payload = np.concatenate([ti, qi], axis=1)
    return dict(kind="se3", vertex_kind=VERTEX_SE3, edge_kind=EDGE_SE3, vertex_ids=ids.astype(np.int32),
                vertex_payload=vert_payload, edge_v0=v0.astype(np.int32), edge_v1=v1.astype(np.int32),
                edge_payload=edge_payload, truth_payload=np.concatenate([t, q], axis=1))


def venice_like(num_cams=871, num_points=530304, mean_extra_obs=1.8, seed=871, pixel_sigma=1.0, fixed_obs=None):
    """Bundle adjustment shaped like Venice (871 cameras / 530k points / ~2M observations).

    Cameras sit on a ring looking inward; every point is seen by k = 2 + Poisson(mean_extra_obs) cameras that
    form a contiguous window of the ring (banded + wrap-around reduced camera matrix). `fixed_obs` forces a
    constant k (config 4 uses k = 10)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Rc = 10.0
    phi = 2 * np.pi * np.arange(num_cams) / num_cams
    C = np.stack([Rc * np.cos(phi), Rc * np.sin(phi), np.zeros(num_cams)], axis=1)
    zc = -C / np.linalg.norm(C, axis=1, keepdims=True)            # optical axis: towards the centre
    xc = np.stack([-np.sin(phi), np.cos(phi), np.zeros(num_cams)], axis=1)
    yc = np.cross(zc, xc)
    Rm = np.stack([xc, yc, zc], axis=2)                            # camera-to-world rotation, columns = axes
    q = _rot_to_quat(Rm)
    fx = fy = 1000.0
    r = 4.0 * np.sqrt(rng.uniform(0, 1, num_points))
    ang = rng.uniform(0, 2 * np.pi, num_points)
    X = np.stack([r * np.cos(ang), r * np.sin(ang), rng.uniform(-2, 2, num_points)], axis=1)
    k = np.full(num_points, fixed_obs) if fixed_obs else 2 + rng.poisson(mean_extra_obs, num_points)
    k = np.minimum(k, num_cams).astype(np.int64)
    start = rng.integers(0, num_cams, num_points)
    pt_of_edge = np.repeat(np.arange(num_points), k)
    offs = np.arange(k.sum()) - np.repeat(np.cumsum(k) - k, k)
    cam_of_edge = (start[pt_of_edge] + offs) % num_cams
    # exact projections + pixel noise
    Rt = np.transpose(Rm, (0, 2, 1))
    pc = np.einsum("eij,ej->ei", Rt[cam_of_edge], X[pt_of_edge] - C[cam_of_edge])
    uv = np.stack([fx * pc[:, 0] / pc[:, 2], fy * pc[:, 1] / pc[:, 2]], axis=1) + rng.normal(0, pixel_sigma, (len(pc), 2))
    # perturbed initial state
    Ci = C + rng.normal(0, 0.02, C.shape)
    dq = np.concatenate([rng.normal(0, 0.001, (num_cams, 3)), np.ones((num_cams, 1))], axis=1)
    dq /= np.linalg.norm(dq, axis=1, keepdims=True)
    qi = _qmul(q, dq)
    Xi = X + rng.normal(0, 0.05, X.shape)
    cam_payload = np.concatenate([Ci, qi, np.tile(np.array([fx, fy, 0.0, 0.0, 0.0]), (num_cams, 1))], axis=1)
    cam_ids = np.arange(num_cams, dtype=np.int32)
    pt_ids = (num_cams + np.arange(num_points)).astype(np.int32)
    return dict(kind="ba", cam_ids=cam_ids, cam_payload=cam_payload, point_ids=pt_ids, point_payload=Xi,
                edge_v0=pt_ids[pt_of_edge], edge_v1=cam_ids[cam_of_edge], edge_payload=uv,
                truth_points=X, truth_cams=np.concatenate([C, q], axis=1))


def expmap_ba(num_cams=30, num_points=600, seed=7, focal_length=1000.0, principal_point=(320.0, 240.0), **kw):
    """The same scene in the SE3-expmap formulation of examples/ba/ba_demo.cpp:95-185: VERTEX_SE3:EXPMAP poses (the file
    holds cam2world: camera centre + camera-to-world quaternion), one PARAMS_CAMERAPARAMETERS (id 0) and
    EDGE_PROJECT_XYZ2UV:EXPMAP observations `paramId u v i00 i01 i11`."""
    p = venice_like(num_cams, num_points, seed=seed, **kw)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    cx, cy = principal_point
    uv = p["edge_payload"] * (focal_length / 1000.0) + np.array([cx, cy])
    n = len(uv)
    w = rng.uniform(0.5, 2.0, (n, 2))                     # non-trivial information matrices (upper triangle i00 i01 i11)
    c = rng.uniform(-0.3, 0.3, n) * np.sqrt(w[:, 0] * w[:, 1])
    pay = np.concatenate([np.zeros((n, 1)), uv, w[:, :1], c[:, None], w[:, 1:]], axis=1)
    return dict(kind="ba_expmap", camera_parameters={0: (focal_length, cx, cy, 0.0)}, cam_ids=p["cam_ids"],
                cam_payload=np.ascontiguousarray(p["cam_payload"][:, :7]), point_ids=p["point_ids"],
                point_payload=p["point_payload"], edge_v0=p["edge_v0"], edge_v1=p["edge_v1"], edge_payload=pay,
                truth_points=p["truth_points"], truth_cams=p["truth_cams"])


def ba_demo(pixel_noise=1.0, outlier_ratio=0.0, seed=1):
    """The scene of the reference's examples/ba/ba_demo.cpp:170-250 in .g2o payload form: 500 points in the box
    [-1.5, 1.5] x [-0.5, 0.5] x [3, 4], 15 cameras with identity rotation at x = 0.04 i - 1 (world -> camera translation,
    as ba_demo sets the estimate), focal length 1000, principal point (320, 240); a point is kept when at least two
    cameras see it inside the 640 x 480 image; pixel noise N(0, pixel_noise), initial points = truth + N(0, 1) per axis;
    with `outlier_ratio` an observation is replaced by a uniform pixel.  The first two poses are meant to be fixed
    (`fixed_ids`).  numpy PCG64 stands in for the reference's std::rand."""
    rng = np.random.Generator(np.random.PCG64(seed))
    truth = np.stack([(rng.uniform(0, 1, 500) - 0.5) * 3, rng.uniform(0, 1, 500) - 0.5, rng.uniform(0, 1, 500) + 3], axis=1)
    f, cx, cy = 1000.0, 320.0, 240.0
    trans = np.stack([np.arange(15) * 0.04 - 1.0, np.zeros(15), np.zeros(15)], axis=1)   # world -> camera, R = I
    # the .g2o vertex line holds cam2world = the inverse: same rotation, translation negated
    cam_payload = np.concatenate([-trans, np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (15, 1))], axis=1)
    pc = truth[:, None, :] + trans[None, :, :]                                          # [point, camera, 3]
    z = np.stack([f * pc[..., 0] / pc[..., 2] + cx, f * pc[..., 1] / pc[..., 2] + cy], axis=-1)
    vis = (z[..., 0] >= 0) & (z[..., 1] >= 0) & (z[..., 0] < 640) & (z[..., 1] < 480)
    keep = vis.sum(1) >= 2
    truth_kept = truth[keep]
    init = truth_kept + rng.normal(0, 1, truth_kept.shape)
    pi, ci = np.nonzero(vis[keep])
    uv = z[keep][pi, ci]
    outlier = rng.uniform(0, 1, len(uv)) < outlier_ratio
    uv = np.where(outlier[:, None], np.stack([rng.uniform(0, 640, len(uv)), rng.uniform(0, 480, len(uv))], axis=1), uv)
    uv = uv + rng.normal(0, pixel_noise, uv.shape)
    inlier = np.ones(len(truth_kept), bool)
    inlier[np.unique(pi[outlier])] = False
    n = len(uv)
    pay = np.concatenate([np.zeros((n, 1)), uv, np.ones((n, 1)), np.zeros((n, 1)), np.ones((n, 1))], axis=1)
    cam_ids = np.arange(15, dtype=np.int32)
    pt_ids = (15 + np.arange(len(truth_kept))).astype(np.int32)
    return dict(kind="ba_expmap", camera_parameters={0: (f, cx, cy, 0.0)}, cam_ids=cam_ids, cam_payload=cam_payload,
                point_ids=pt_ids, point_payload=init, edge_v0=pt_ids[pi], edge_v1=cam_ids[ci], edge_payload=pay,
                truth_points=truth_kept, inlier_points=inlier, fixed_ids=np.array([0, 1], dtype=np.int32))


def _rot_to_quat(R):
    """batched rotation matrix -> quaternion (x y z w), w >= 0"""
    m = R
    tr = m[:, 0, 0] + m[:, 1, 1] + m[:, 2, 2]
    q = np.zeros((len(R), 4))
    for i in range(len(R)):
        M = m[i]
        if tr[i] > 0:
            s = np.sqrt(tr[i] + 1.0) * 2
            q[i] = [(M[2, 1] - M[1, 2]) / s, (M[0, 2] - M[2, 0]) / s, (M[1, 0] - M[0, 1]) / s, 0.25 * s]
        else:
            a = int(np.argmax([M[0, 0], M[1, 1], M[2, 2]]))
            b, c = (a + 1) % 3, (a + 2) % 3
            s = np.sqrt(1.0 + M[a, a] - M[b, b] - M[c, c]) * 2
            v = np.zeros(4)
            v[a] = 0.25 * s
            v[b] = (M[b, a] + M[a, b]) / s
            v[c] = (M[c, a] + M[a, c]) / s
            v[3] = (M[c, b] - M[b, c]) / s
            q[i] = v
        if q[i, 3] < 0:
            q[i] = -q[i]
    return q / np.linalg.norm(q, axis=1, keepdims=True)


def _interleaved_ids(n_poses, n_lm):
    """vertex ids with poses and landmarks interleaved (the index mapping sorts by id: sparse_optimizer.cpp:166-190)"""
    pose_ids = 3 * np.arange(n_poses)
    lm_ids = 3 * (np.arange(n_lm) % n_poses) + 1 + (np.arange(n_lm) // n_poses) * 3 * n_poses
    lm_ids = np.where(np.arange(n_lm) < n_poses, 3 * np.arange(n_lm) + 1, 3 * n_poses + np.arange(n_lm))
    return pose_ids.astype(np.int32), lm_ids.astype(np.int32)


def landmark_slam_2d(n_poses=120, n_landmarks=60, seed=32, max_range=6.0, sigma_odo=(0.02, 0.02, 0.01), sigma_obs=0.05,
                     odometry=True, growth=0.0):
    """2D landmark SLAM: a robot drives laps on a slowly drifting circle, odometry EdgeSE2 between consecutive poses
    (+ one loop closure per lap) and EdgeSE2PointXY sightings of the landmarks in range (types/slam2d/edge_se2_pointxy.h).
    Initial guess: integrated odometry; landmarks from their first sighting.  growth > 0: the circle widens by that much
    per pose (a spiral: the robot explores, every landmark is seen from a bounded number of poses)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = np.arange(n_poses)
    per_lap = 40
    ang = 2 * np.pi * k / per_lap
    rad = 8.0 + 0.5 * np.sin(0.37 * k) + growth * k
    x, y = rad * np.cos(ang), rad * np.sin(ang)
    th = ang + np.pi / 2 + 0.1 * np.sin(0.61 * k)
    th = (th + np.pi) % (2 * np.pi) - np.pi
    la = rng.uniform(0, 2 * np.pi, n_landmarks)
    lr = rng.uniform(3.0, 13.0, n_landmarks) if growth == 0 else np.sqrt(rng.uniform(0.0, 1.0, n_landmarks)) * (rad.max() + 5.0)
    lm = np.stack([lr * np.cos(la), lr * np.sin(la)], axis=1)

    def rel(i, j):  # x_i^-1 * x_j
        c, s = np.cos(th[i]), np.sin(th[i])
        dx, dy = x[j] - x[i], y[j] - y[i]
        dth = (th[j] - th[i] + np.pi) % (2 * np.pi) - np.pi
        return np.stack([c * dx + s * dy, -s * dx + c * dy, dth], axis=-1)
    o0 = np.concatenate([k[:-1], k[per_lap:]])
    o1 = np.concatenate([k[1:], k[:-per_lap]])
    so = np.asarray(sigma_odo)
    zo = rel(o0, o1) + rng.normal(0, 1, (len(o0), 3)) * so
    zo[:, 2] = (zo[:, 2] + np.pi) % (2 * np.pi) - np.pi
    io = np.diag(1.0 / so ** 2)
    iu = np.array([io[i, j] for i in range(3) for j in range(i, 3)])
    odo_payload = np.concatenate([zo, np.tile(iu, (len(o0), 1))], axis=1)
    # sightings
    d = lm[None, :, :] - np.stack([x, y], axis=1)[:, None, :]
    seen = np.linalg.norm(d, axis=2) < max_range
    pi_, li_ = np.nonzero(seen)
    c, s = np.cos(th[pi_]), np.sin(th[pi_])
    dl = d[pi_, li_]
    zl = np.stack([c * dl[:, 0] + s * dl[:, 1], -s * dl[:, 0] + c * dl[:, 1]], axis=1) + rng.normal(0, sigma_obs, (len(pi_), 2))
    w = 1.0 / sigma_obs ** 2
    obs_payload = np.concatenate([zl, np.tile([w, 0.1 * w, 1.3 * w], (len(pi_), 1))], axis=1)
    used = np.unique(li_)
    remap = -np.ones(n_landmarks, int)
    remap[used] = np.arange(len(used))
    # initial guess
    pe = np.zeros((n_poses, 3))
    pe[0] = [x[0], y[0], th[0]]
    for i in range(1, n_poses):
        c0, s0 = np.cos(pe[i - 1, 2]), np.sin(pe[i - 1, 2])
        z = zo[i - 1]
        pe[i] = [pe[i - 1, 0] + c0 * z[0] - s0 * z[1], pe[i - 1, 1] + s0 * z[0] + c0 * z[1],
                 (pe[i - 1, 2] + z[2] + np.pi) % (2 * np.pi) - np.pi]
    le = np.zeros((len(used), 2))
    first = {}
    for q, (p, l) in enumerate(zip(pi_, li_)):
        first.setdefault(int(l), q)
    for l, q in first.items():
        p = pi_[q]
        c0, s0 = np.cos(pe[p, 2]), np.sin(pe[p, 2])
        le[remap[l]] = [pe[p, 0] + c0 * zl[q, 0] - s0 * zl[q, 1], pe[p, 1] + s0 * zl[q, 0] + c0 * zl[q, 1]]
    pose_ids, lm_ids = _interleaved_ids(n_poses, len(used))
    if not odometry:
        o0, o1, odo_payload = o0[:0], o1[:0], odo_payload[:0]
    return dict(kind="slam2d", pose_ids=pose_ids, pose_payload=pe, lm_ids=lm_ids, lm_payload=le,
                odo_v0=pose_ids[o0], odo_v1=pose_ids[o1], odo_payload=odo_payload,
                obs_v0=pose_ids[pi_], obs_v1=lm_ids[remap[li_]], obs_payload=obs_payload,
                truth_poses=np.stack([x, y, th], axis=1), truth_landmarks=lm[used])


def landmark_slam_3d(n_poses=80, n_landmarks=50, seed=33, max_range=7.0, sigma_t=0.02, sigma_r=0.01, sigma_obs=0.04,
                     offset=(0.1, -0.05, 0.2, 0.1, 0.2, -0.1, 0.9)):
    """3D landmark SLAM: SE3 poses on a helix with EdgeSE3 odometry (+ closures to the previous turn) and EdgeSE3PointXYZ
    sightings (types/slam3d/edge_se3_pointxyz.cpp) through one ParameterSE3Offset (sensor pose on the robot, x y z qx qy qz qw)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    k = np.arange(n_poses)
    per_turn = 20
    ang = 2 * np.pi * k / per_turn
    t = np.stack([6.0 * np.cos(ang), 6.0 * np.sin(ang), 0.15 * k], axis=1)
    qz = np.stack([0 * ang, 0 * ang, np.sin((ang + np.pi / 2) / 2), np.cos((ang + np.pi / 2) / 2)], axis=-1)
    tilt = 0.2 * np.sin(0.3 * k)
    qx = np.stack([np.sin(tilt / 2), 0 * tilt, 0 * tilt, np.cos(tilt / 2)], axis=-1)
    q = _qmul(qz, qx)
    lm = np.stack([rng.uniform(-9, 9, n_landmarks), rng.uniform(-9, 9, n_landmarks), rng.uniform(-1, 0.15 * n_poses + 1, n_landmarks)], axis=1)
    o0 = np.concatenate([k[:-1], k[per_turn:]])
    o1 = np.concatenate([k[1:], k[:-per_turn]])
    qrel = _qmul(_qconj(q[o0]), q[o1])
    trel = _qrot(_qconj(q[o0]), t[o1] - t[o0])
    E = len(o0)
    v = rng.normal(0.0, sigma_r, (E, 3))
    qn = np.concatenate([v, np.sqrt(np.maximum(0.0, 1.0 - (v ** 2).sum(1)))[:, None]], axis=1)
    qmeas = _qmul(qrel, qn)
    qmeas /= np.linalg.norm(qmeas, axis=1, keepdims=True)
    tmeas = trel + rng.normal(0.0, sigma_t, (E, 3))
    info = np.zeros((6, 6))
    info[:3, :3] = np.eye(3) / sigma_t ** 2
    info[3:, 3:] = np.eye(3) / sigma_r ** 2
    iu = np.array([info[i, j] for i in range(6) for j in range(i, 6)])
    odo_payload = np.concatenate([tmeas, qmeas, np.tile(iu, (E, 1))], axis=1)
    off = np.asarray(offset, float)
    oq = off[3:] / np.linalg.norm(off[3:])
    # sensor frame: n2w = X * offset
    sq = _qmul(q, np.tile(oq, (n_poses, 1)))
    st = t + _qrot(q, np.tile(off[:3], (n_poses, 1)))
    d = lm[None, :, :] - st[:, None, :]
    seen = np.linalg.norm(d, axis=2) < max_range
    pi_, li_ = np.nonzero(seen)
    zl = _qrot(_qconj(sq[pi_]), d[pi_, li_]) + rng.normal(0, sigma_obs, (len(pi_), 3))
    w = 1.0 / sigma_obs ** 2
    iw = np.array([w, 0.05 * w, 0.0, 1.2 * w, -0.1 * w, 0.9 * w])
    obs_payload = np.concatenate([np.zeros((len(pi_), 1)), zl, np.tile(iw, (len(pi_), 1))], axis=1)  # paramId 0
    used = np.unique(li_)
    remap = -np.ones(n_landmarks, int)
    remap[used] = np.arange(len(used))
    qo = np.concatenate([q[:1], qmeas[:n_poses - 1]], axis=0)
    to = np.concatenate([t[:1], tmeas[:n_poses - 1]], axis=0)
    qi, ti = _prefix_compose(qo, to)
    sqi = _qmul(qi, np.tile(oq, (n_poses, 1)))
    sti = ti + _qrot(qi, np.tile(off[:3], (n_poses, 1)))
    le = np.zeros((len(used), 3))
    first = {}
    for qq, l in enumerate(li_):
        first.setdefault(int(l), qq)
    for l, qq in first.items():
        p = pi_[qq]
        le[remap[l]] = sti[p] + _qrot(sqi[p][None, :], zl[qq][None, :])[0]
    pose_ids, lm_ids = _interleaved_ids(n_poses, len(used))
    return dict(kind="slam3d", pose_ids=pose_ids, pose_payload=np.concatenate([ti, qi], axis=1), lm_ids=lm_ids, lm_payload=le,
                odo_v0=pose_ids[o0], odo_v1=pose_ids[o1], odo_payload=odo_payload,
                obs_v0=pose_ids[pi_], obs_v1=lm_ids[remap[li_]], obs_payload=obs_payload,
                offsets={0: np.concatenate([off[:3], oq])},
                truth_poses=np.concatenate([t, q], axis=1), truth_landmarks=lm[used])


def feed(problem, target):
    """push a generated problem into anything with add_vertices/add_edges (product SparseOptimizer or the
    tests' oracle wrapper)"""
    if problem["kind"] == "se3":
        target.add_vertices(VERTEX_SE3, problem["vertex_ids"], problem["vertex_payload"])
        target.add_edges(EDGE_SE3, problem["edge_v0"], problem["edge_v1"], problem["edge_payload"])
    elif problem["kind"] in ("slam2d", "slam3d"):
        three_d = problem["kind"] == "slam3d"
        for pid, off in problem.get("offsets", {}).items():
            target.add_se3_offset(pid, off)
        target.add_vertices(VERTEX_SE3 if three_d else VERTEX_SE2, problem["pose_ids"], problem["pose_payload"])
        target.add_vertices(VERTEX_XYZ if three_d else VERTEX_XY, problem["lm_ids"], problem["lm_payload"])
        if len(problem["odo_v0"]):
            target.add_edges(EDGE_SE3 if three_d else EDGE_SE2, problem["odo_v0"], problem["odo_v1"], problem["odo_payload"])
        target.add_edges(EDGE_SE3_XYZ if three_d else EDGE_SE2_XY, problem["obs_v0"], problem["obs_v1"], problem["obs_payload"])
    elif problem["kind"] == "ba_expmap":
        for pid, par in problem["camera_parameters"].items():
            target.add_camera_parameters(pid, *par)
        target.add_vertices(VERTEX_SE3_EXPMAP, problem["cam_ids"], problem["cam_payload"])
        target.add_vertices(VERTEX_XYZ, problem["point_ids"], problem["point_payload"])
        target.add_edges(EDGE_XYZ2UV, problem["edge_v0"], problem["edge_v1"], problem["edge_payload"])
    else:
        target.add_vertices(VERTEX_CAM, problem["cam_ids"], problem["cam_payload"])
        target.add_vertices(VERTEX_XYZ, problem["point_ids"], problem["point_payload"])
        target.add_edges(EDGE_P2MC, problem["edge_v0"], problem["edge_v1"], problem["edge_payload"])


def write_g2o(problem, path):
    """text form of a generated problem (the on-disk format of core/optimizable_graph.cpp:356-569)"""
    with open(path, "w") as f:
        if problem["kind"] == "se3":
            for i, p in zip(problem["vertex_ids"], problem["vertex_payload"]):
                f.write("VERTEX_SE3:QUAT %d %s\n" % (i, " ".join(repr(float(x)) for x in p)))
            for a, b, p in zip(problem["edge_v0"], problem["edge_v1"], problem["edge_payload"]):
                f.write("EDGE_SE3:QUAT %d %d %s\n" % (a, b, " ".join(repr(float(x)) for x in p)))
        elif problem["kind"] in ("slam2d", "slam3d"):
            three_d = problem["kind"] == "slam3d"
            num = lambda p: " ".join(repr(float(x)) for x in p)
            for pid, off in problem.get("offsets", {}).items():
                f.write("PARAMS_SE3OFFSET %d %s\n" % (pid, num(off)))
            for i, p in zip(problem["pose_ids"], problem["pose_payload"]):
                f.write("%s %d %s\n" % ("VERTEX_SE3:QUAT" if three_d else "VERTEX_SE2", i, num(p)))
            for i, p in zip(problem["lm_ids"], problem["lm_payload"]):
                f.write("%s %d %s\n" % ("VERTEX_TRACKXYZ" if three_d else "VERTEX_XY", i, num(p)))
            for a, b, p in zip(problem["odo_v0"], problem["odo_v1"], problem["odo_payload"]):
                f.write("%s %d %d %s\n" % ("EDGE_SE3:QUAT" if three_d else "EDGE_SE2", a, b, num(p)))
            for a, b, p in zip(problem["obs_v0"], problem["obs_v1"], problem["obs_payload"]):
                if three_d:
                    f.write("EDGE_SE3_TRACKXYZ %d %d %d %s\n" % (a, b, int(p[0]), num(p[1:])))
                else:
                    f.write("EDGE_SE2_XY %d %d %s\n" % (a, b, num(p)))
        elif problem["kind"] == "ba_expmap":
            for pid, par in problem["camera_parameters"].items():
                f.write("PARAMS_CAMERAPARAMETERS %d %s\n" % (pid, " ".join(repr(float(x)) for x in par)))
            for i, p in zip(problem["cam_ids"], problem["cam_payload"]):
                f.write("VERTEX_SE3:EXPMAP %d %s\n" % (i, " ".join(repr(float(x)) for x in p)))
            for i, p in zip(problem["point_ids"], problem["point_payload"]):
                f.write("VERTEX_XYZ %d %s\n" % (i, " ".join(repr(float(x)) for x in p)))
            for a, b, p in zip(problem["edge_v0"], problem["edge_v1"], problem["edge_payload"]):
                f.write("EDGE_PROJECT_XYZ2UV:EXPMAP %d %d %d %s\n" % (a, b, int(p[0]), " ".join(repr(float(x)) for x in p[1:])))
        else:
            for i, p in zip(problem["cam_ids"], problem["cam_payload"]):
                f.write("VERTEX_CAM %d %s\n" % (i, " ".join(repr(float(x)) for x in p)))
            for i, p in zip(problem["point_ids"], problem["point_payload"]):
                f.write("VERTEX_XYZ %d %s\n" % (i, " ".join(repr(float(x)) for x in p)))
            for a, b, p in zip(problem["edge_v0"], problem["edge_v1"], problem["edge_payload"]):
                f.write("EDGE_PROJECT_P2MC %d %d %s\n" % (a, b, " ".join(repr(float(x)) for x in p)))
