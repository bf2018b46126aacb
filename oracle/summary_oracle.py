"""NumPy restatement of the first consumer of the pasted masks: SummaryOutput with
CrackToInstance, CalculateInstanceSize and IncludeMyRoad.
TEST INFRASTRUCTURE — see oracle/__init__.py ("parity unpinned").

Every function cites the reference file:line it follows (paths relative to /root/reference).

Floating-point definition.  The reference sums float32 products with tf.reduce_sum (Eigen tree
order, not reproducible without TensorFlow) and solves a 2x2 least-squares system with
tf.linalg.inv in float32 (condition number ~1e5..1e6 for 512..1080 rows, i.e. percent-level noise in
TF itself).  The restatement therefore DEFINES every reduction as the exact sum of the float32
terms accumulated in float64 and rounded to float32 once, and the line fit as the closed-form
normal-equation solution evaluated in float64 on exact integer moments, rounded to float32 once.
Element-wise arithmetic (unit**2, y*theta0+theta1, clip, divide, compare) stays float32, one
rounding per operation, in the reference's order.
"""
import numpy as np

F32 = np.float32
F64 = np.float64


# ------------------------------------------------------------------ CrackToInstance
def crack_to_instance(crack, crack_id=5):
    """engine/layers/misc.py:515-537.  crack: seg_outs[..., 2], int [B,PH,PW].
    Returns (crack_det_outs [B,1,6] int32, crack_seg_outs [B,1,PH,PW] float32).  The box is the
    bounding box of the non-zero pixels of the WHOLE batch (tf.where over [B,PH,PW], min/max over
    axis 0); the class id is the literal 5 (the layer ignores its crack_id argument)."""
    crack = np.asarray(crack)
    B = crack.shape[0]
    idx = np.argwhere(crack != 0)
    if idx.size == 0:
        idx = np.zeros((1, 3), dtype=np.int64)                    # :517-519
    ymin, xmin = idx.min(axis=0)[1:]
    ymax, xmax = idx.max(axis=0)[1:]
    height = np.int32(ymax - ymin)
    width = np.int32(xmax - xmin)
    cy = np.int32(ymin) + np.int32(np.trunc(height / 2))          # int / 2 -> float64 -> cast
    cx = np.int32(xmin) + np.int32(np.trunc(width / 2))
    class_id = np.int32(5)
    conf = np.int32(np.clip(np.int64(100) * height * width, 0, 100))
    row = np.array([cx, cy, width, height, class_id, conf], dtype=np.int32)
    det = np.tile(row[None, None, :], (B, 1, 1))
    seg = crack[:, None].astype(F32)
    return det, seg


# ------------------------------------------------------------- CalculateInstanceSize
def fit_line(pos):
    """_calculate_theta, misc.py:706-718: least squares x = theta0*y + theta1 through the points
    pos [n,2] = (y, x); zeros when det(X^T X) <= 0.  Closed form on exact integer moments in
    float64 (module docstring), rounded to float32."""
    pos = np.asarray(pos, dtype=np.int64).reshape(-1, 2)
    n = F64(pos.shape[0])
    sy = F64(pos[:, 0].sum())
    syy = F64((pos[:, 0] * pos[:, 0]).sum())
    sx = F64(pos[:, 1].sum())
    sxy = F64((pos[:, 0] * pos[:, 1]).sum())
    det = syy * n - sy * sy
    if not det > 0:
        return F32(0), F32(0)
    t0 = (n * sxy - sy * sx) / det
    t1 = (syy * sx - sy * sxy) / det
    return F32(t0), F32(t1)


def fit_line_f32(pos, order="blas"):
    """What TensorFlow itself computes in _calculate_theta (misc.py:706-718), emulated in float32 to QUANTIFY
    how far the float64 contract above sits from the reference's own arithmetic (it is not the contract):
    xs = [y, 1] and ys = x cast to float32, X^T X and X^T y as float32 matrix products, tf.linalg.det's sign
    test, tf.linalg.inv as a float32 partial-pivoting LU (LAPACK sgetrf/sgetri here, Eigen PartialPivLU in TF),
    then the float32 product inv @ (X^T y).  The accumulation order inside TF's matmul is an implementation
    detail, so two orders are offered: "blas" (NumPy's sgemm) and "sequential" (one float32 add per row)."""
    pos = np.asarray(pos, dtype=np.int64).reshape(-1, 2)
    if pos.shape[0] == 0:
        return F32(0), F32(0)
    xs = np.stack([pos[:, 0].astype(F32), np.ones(pos.shape[0], F32)], axis=1)       # [n,2]
    ys = pos[:, 1].astype(F32)[:, None]                                             # [n,1]
    if order == "blas":
        x_mat = (xs.T @ xs).astype(F32)
        x_y = (xs.T @ ys).astype(F32)
    else:
        def seq(terms):
            acc = F32(0)
            for t in terms:
                acc = F32(acc + t)
            return acc
        y = xs[:, 0]
        x_mat = np.array([[seq(y * y), seq(y)], [seq(y), seq(np.ones_like(y))]], dtype=F32)
        x_y = np.array([[seq(y * ys[:, 0])], [seq(ys[:, 0])]], dtype=F32)
    det = F32(F32(x_mat[0, 0] * x_mat[1, 1]) - F32(x_mat[0, 1] * x_mat[1, 0]))
    if not det > 0:
        return F32(0), F32(0)
    inv = np.linalg.inv(x_mat).astype(F32)                                         # float32 LU (sgesv)
    theta = (inv @ x_y).astype(F32)
    return F32(theta[0, 0]), F32(theta[1, 0])


def road_marginals(road):
    """_calculate_marginal_x_by_y_axis, misc.py:680-704.  road: int [PH,PW].  tf.segment_min/max
    over the sorted row ids give one (x_min, x_max) per row up to the last road row, 0 for rows
    without road pixels; rows with x_min == x_max are dropped, then 15 % (at least one row) at
    both ends."""
    road = np.asarray(road)
    ys, xs = np.nonzero(road > 0)
    if ys.size == 0:
        e = np.zeros((0, 2), dtype=np.int64)
        return e, e
    L = int(ys.max()) + 1
    x_min = np.zeros(L, dtype=np.int64)
    x_max = np.zeros(L, dtype=np.int64)
    for y in np.unique(ys):
        r = xs[ys == y]
        x_min[y], x_max[y] = r.min(), r.max()
    keep = np.nonzero(x_min != x_max)[0]
    left = np.stack([keep, x_min[keep]], axis=1)
    right = np.stack([keep, x_max[keep]], axis=1)
    valid = F32(left.shape[0])
    drop = int(max(1, np.int32(valid * F32(0.15))))               # clip(int(n*0.15), 1, 2**31)
    return left[drop:left.shape[0] - drop], right[drop:right.shape[0] - drop]


def road_unit_lengths(road, default_road_size=3.25, fit=None):
    """_calculate_road_size_by_vertical_per_batch, misc.py:660-678 -> float32 [PH]: metres per
    pixel on every frame row from the fitted left/right road borders.  `fit` swaps the line fit
    (default: the float64 contract `fit_line`; `fit_line_f32` to measure the distance to TF's float32)."""
    fit = fit or fit_line
    road = np.asarray(road)
    left, right = road_marginals(road)
    l0, l1 = fit(left)
    r0, r1 = fit(right)
    y = np.arange(road.shape[0], dtype=F32)
    pred_left = y * l0 + l1
    pred_right = y * r0 + r1
    width = np.maximum(pred_right - pred_left, F32(1.0)).astype(F32)     # clip_by_value(.., 1, inf)
    return (F32(default_road_size) / width).astype(F32)


def instance_reductions(seg_outs, masks, default_road_size=3.25, threshold=0.1, road_channel=1):
    """The five per-instance reductions over [B,M,PH,PW] float masks:
    pixel_counts (misc.py:573-574), instance/horizontal/vertical size (CalculateInstanceSize.call,
    misc.py:633-658) and include_my_road (IncludeMyRoad.call, misc.py:603-617).
    Returns float32 [B,M,5] in that order."""
    seg_outs = np.asarray(seg_outs)
    masks = np.asarray(masks, dtype=F32)
    B, M, PH, PW = masks.shape
    out = np.zeros((B, M, 5), dtype=F32)
    for b in range(B):
        road = seg_outs[b, :, :, road_channel]
        unit = road_unit_lengths(road, default_road_size)              # [PH] f32
        unit2 = (unit * unit).astype(F32)                              # unit ** 2
        my_road = road.astype(F32) > 0.5
        for j in range(M):
            m = masks[b, j].astype(F64)
            binm = masks[b, j] > 0.5
            out[b, j, 0] = F32(m.sum())
            out[b, j, 1] = F32((unit2.astype(F64)[:, None] * m).sum())
            out[b, j, 2] = F32((unit.astype(F64)[:, None] * m).sum(axis=0).max())
            out[b, j, 3] = F32((unit.astype(F64) * binm.any(axis=1)).sum())
            inter = F32(np.count_nonzero(my_road & binm))
            area = F32(np.count_nonzero(binm))
            ioi = inter / (area + F32(1e-5))
            out[b, j, 4] = F32(1.0) if ioi > F32(threshold) else F32(0.0)
    return out


def calculate_instance_size(seg_outs, pad_ins_outs, default_road_size=3.25):
    """CalculateInstanceSize.call -> [B,M,3] (instance, horizontal, vertical)."""
    return instance_reductions(seg_outs, pad_ins_outs, default_road_size)[..., 1:4]


def include_my_road(seg_outs, crop_ins_outs, threshold=0.1):
    """IncludeMyRoad.call -> [B,M] float 0/1."""
    return instance_reductions(seg_outs, crop_ins_outs, threshold=threshold)[..., 4]


# --------------------------------------------------------------------- SummaryOutput
def summary_output(det_outs, seg_outs, crop_ins_outs, default_road_size=3.25):
    """SummaryOutput.call, misc.py:554-583 -> float32 [B,M',11]:
    (class, cx, cy, w, h, conf, pixel_counts, instance_size, horizontal_size, vertical_size,
    include_my_road); M' = M + 1 when the batch holds a crack region with positive area."""
    det_outs = np.asarray(det_outs, dtype=np.int32)
    seg_outs = np.asarray(seg_outs)
    masks = np.asarray(crop_ins_outs, dtype=F32)
    crack_det, crack_seg = crack_to_instance(seg_outs[..., 2])
    if np.all(crack_det[..., -1] > 0):                                # :562-568
        det_outs = np.concatenate([det_outs, crack_det], axis=1)
        masks = np.concatenate([masks, crack_seg], axis=1)
    d = det_outs[..., :6].astype(F32)
    red = instance_reductions(seg_outs, masks, default_road_size)
    cols = [d[..., 4], d[..., 0], d[..., 1], d[..., 2], d[..., 3], d[..., 5],
            red[..., 0], red[..., 1], red[..., 2], red[..., 3], red[..., 4]]
    return np.stack(cols, axis=-1).astype(F32)
