"""numpy restatements of the reference's host-side decode path and NMS (test oracle)."""
import numpy as np

MPII_FLIP_PAIRS = [[0, 5], [1, 4], [2, 3], [10, 15], [11, 14], [12, 13]]  # lib/dataset/mpii.py:32


def get_max_preds(batch_heatmaps):
    """lib/core/inference.py:18-46: flat arg-max per (b,j) (first maximum), x = idx % W, y = floor(idx / W),
    zeroed where the maximum is <= 0."""
    B, J, H, W = batch_heatmaps.shape
    flat = batch_heatmaps.reshape(B, J, -1)
    idx = np.argmax(flat, 2).reshape(B, J, 1)
    maxvals = np.amax(flat, 2).reshape(B, J, 1)
    preds = np.tile(idx, (1, 1, 2)).astype(np.float32)
    preds[:, :, 0] = preds[:, :, 0] % W
    preds[:, :, 1] = np.floor(preds[:, :, 1] / W)
    preds *= np.tile(np.greater(maxvals, 0.0), (1, 1, 2)).astype(np.float32)
    return preds, maxvals


def flip_back(output_flipped, matched_parts):
    """lib/utils/transforms.py:15-29: reverse W, swap left/right joint channels."""
    out = output_flipped[:, :, :, ::-1].copy()
    for a, b in matched_parts:
        tmp = out[:, a].copy()
        out[:, a] = out[:, b]
        out[:, b] = tmp
    return out


def flip_test_merge(output, output_flipped, matched_parts, shift_heatmap=True):
    """lib/core/function.py:224-240: flip_back, optional 1-px right shift, average."""
    f = flip_back(output_flipped, matched_parts)
    if shift_heatmap:
        f[:, :, :, 1:] = f.copy()[:, :, :, 0:-1]
    return (output + f) * np.float32(0.5)


def nms(dets, thresh):
    """lib/nms/nms.py:35-72 greedy IoU NMS ("+1" pixel convention, keep ovr <= thresh)."""
    if dets.shape[0] == 0:
        return []
    x1, y1, x2, y2, scores = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3], dets[:, 4]
    areas = (x2 - x1 + 1) * (y2 - y1 + 1)
    order = scores.argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        xx1 = np.maximum(x1[i], x1[order[1:]])
        yy1 = np.maximum(y1[i], y1[order[1:]])
        xx2 = np.minimum(x2[i], x2[order[1:]])
        yy2 = np.minimum(y2[i], y2[order[1:]])
        w = np.maximum(0.0, xx2 - xx1 + 1)
        h = np.maximum(0.0, yy2 - yy1 + 1)
        inter = w * h
        ovr = inter / (areas[i] + areas[order[1:]] - inter)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


# ------------------------------------------------------------------------------------------------------
# decode tail: get_final_preds / accuracy / OKS-NMS / Gaussian targets (pinned by tests/golden/decode2.npz)
# ------------------------------------------------------------------------------------------------------
def _affine_from_3pts(src, dst):
    """What cv2.getAffineTransform(src, dst) computes (lib/utils/transforms.py:84-87): the 2x3 matrix M with
    M @ [x, y, 1]^T = dst for the three point pairs, solved in float64 from float32 inputs."""
    a = np.concatenate([np.asarray(src, np.float32).astype(np.float64), np.ones((3, 1))], 1)
    return np.linalg.solve(a, np.asarray(dst, np.float32).astype(np.float64)).T


def get_affine_transform(center, scale, rot, output_size, inv=0):
    """lib/utils/transforms.py:57-89 (shift = 0). Only scale[0] (the width) enters: the reference derives both
    directions from src_w / dst_w."""
    scale_tmp = np.asarray(scale, np.float32) * 200.0
    src_w, dst_w, dst_h = scale_tmp[0], output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    src_dir = np.array([0 * cs - (src_w * -0.5) * sn, 0 * sn + (src_w * -0.5) * cs])      # get_dir, :105-112
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = np.asarray(center, np.float32)
    src[1] = np.asarray(center, np.float32) + src_dir
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    for p in (src, dst):                                                                  # get_3rd_point, :100-102
        d = p[0] - p[1]
        p[2] = p[1] + np.array([-d[1], d[0]], np.float32)
    return _affine_from_3pts(dst, src) if inv else _affine_from_3pts(src, dst)


def transform_preds(coords, center, scale, output_size):
    """lib/utils/transforms.py:49-54: heat-map coordinates -> image coordinates through the inverse affine."""
    t = get_affine_transform(center, scale, 0, output_size, inv=1)
    out = np.zeros(coords.shape)
    for p in range(coords.shape[0]):
        out[p, 0:2] = t @ np.array([coords[p, 0], coords[p, 1], 1.0])
    return out


def get_final_preds(post_process, batch_heatmaps, center, scale):
    """lib/core/inference.py:49-79: arg-max, quarter-pixel nudge toward the higher neighbour when the arg-max is
    strictly inside (1, W-1) x (1, H-1), then transform_preds per sample."""
    coords, maxvals = get_max_preds(batch_heatmaps)
    H, W = batch_heatmaps.shape[2], batch_heatmaps.shape[3]
    if post_process:
        for n in range(coords.shape[0]):
            for p in range(coords.shape[1]):
                hm = batch_heatmaps[n][p]
                px = int(np.floor(coords[n][p][0] + 0.5))
                py = int(np.floor(coords[n][p][1] + 0.5))
                if 1 < px < W - 1 and 1 < py < H - 1:
                    diff = np.array([hm[py][px + 1] - hm[py][px - 1], hm[py + 1][px] - hm[py - 1][px]])
                    coords[n][p] += np.sign(diff) * .25
    preds = coords.copy()
    for i in range(coords.shape[0]):
        preds[i] = transform_preds(coords[i], center[i], scale[i], [W, H])
    return preds, maxvals


def accuracy(output, target, thr=0.5):
    """lib/core/evaluate.py:16-71 (hm_type 'gaussian'): PCK of arg-max(output) against arg-max(target), distances
    normalised by (h, w) / 10; joints whose target arg-max is not > 1 in both coordinates are ignored (-1)."""
    pred, _ = get_max_preds(output)
    tgt, _ = get_max_preds(target)
    h, w = output.shape[2], output.shape[3]
    norm = np.ones((pred.shape[0], 2)) * np.array([h, w]) / 10
    B, J = pred.shape[:2]
    dists = np.zeros((J, B))
    pf, tf = pred.astype(np.float32), tgt.astype(np.float32)
    for n in range(B):
        for c in range(J):
            if tf[n, c, 0] > 1 and tf[n, c, 1] > 1:
                dists[c, n] = np.linalg.norm(pf[n, c] / norm[n] - tf[n, c] / norm[n])
            else:
                dists[c, n] = -1
    acc = np.zeros(J + 1)
    avg, cnt = 0.0, 0
    for c in range(J):
        cal = dists[c] != -1
        acc[c + 1] = (dists[c][cal] < thr).sum() * 1.0 / cal.sum() if cal.sum() > 0 else -1
        if acc[c + 1] >= 0:
            avg += acc[c + 1]
            cnt += 1
    avg = avg / cnt if cnt != 0 else 0
    if cnt != 0:
        acc[0] = avg
    return acc, avg, cnt, pred


COCO_SIGMAS = np.array([.26, .25, .25, .35, .35, .79, .79, .72, .72, .62, .62, 1.07, 1.07, .87, .87, .89, .89]) / 10.0


def oks_iou(g, d, a_g, a_d, sigmas=None, in_vis_thre=None):
    """lib/nms/nms.py:75-96. Quirk kept: `list(vg > t) and list(vd > t)` is Python's `and` of two non-empty lists,
    i.e. just the SECOND list -- the visibility mask is the candidate's alone."""
    sig = COCO_SIGMAS if sigmas is None else np.asarray(sigmas)
    vars_ = (sig * 2) ** 2
    xg, yg = g[0::3], g[1::3]
    ious = np.zeros(d.shape[0])
    for n_d in range(d.shape[0]):
        dx = d[n_d, 0::3] - xg
        dy = d[n_d, 1::3] - yg
        e = (dx ** 2 + dy ** 2) / vars_ / ((a_g + a_d[n_d]) / 2 + np.spacing(1)) / 2
        if in_vis_thre is not None:
            e = e[d[n_d, 2::3] > in_vis_thre]
        ious[n_d] = np.sum(np.exp(-e)) / e.shape[0] if e.shape[0] != 0 else 0.0
    return ious


def oks_nms(kpts, areas, scores, thresh, sigmas=None, in_vis_thre=None):
    """lib/nms/nms.py:99-124 on arrays: kpts [N,J,3], areas [N], scores [N] -> kept indices, best first."""
    if len(scores) == 0:
        return []
    flat = np.asarray(kpts).reshape(len(scores), -1)
    order = np.asarray(scores).argsort()[::-1]
    keep = []
    while order.size > 0:
        i = order[0]
        keep.append(int(i))
        ovr = oks_iou(flat[i], flat[order[1:]], areas[i], np.asarray(areas)[order[1:]], sigmas, in_vis_thre)
        order = order[np.where(ovr <= thresh)[0] + 1]
    return keep


def rescore(kpts, box_scores, in_vis_thre):
    """lib/dataset/coco.py:346-357: person score = box score x mean confidence of the joints above in_vis_thre."""
    out = np.zeros(len(box_scores))
    for i in range(len(box_scores)):
        ks, vn = 0.0, 0
        for t_s in kpts[i, :, 2]:            # sequential sum, like the reference loop
            if t_s > in_vis_thre:
                # float64 accumulation: `0 + np.float32` promoted to float64 under the numpy (< 2) the reference ran on;
                # numpy >= 2 (NEP 50) would keep float32 here -- the golden vectors use float64 key points, where both agree
                ks, vn = ks + float(t_s), vn + 1
        if vn != 0:
            ks = ks / vn
        out[i] = ks * box_scores[i]
    return out


def generate_target(joints, joints_vis, image_size, heatmap_size, sigma=2, joints_weight=None):
    """lib/dataset/JointsDataset.py:233-289 for one sample: joints [J,3], joints_vis [J,3]; image_size / heatmap_size
    are (w, h). Returns (target [J,h,w], target_weight [J,1])."""
    J = joints.shape[0]
    W, H = int(heatmap_size[0]), int(heatmap_size[1])
    tw = np.ones((J, 1), np.float32)
    tw[:, 0] = joints_vis[:, 0]
    target = np.zeros((J, H, W), np.float32)
    tmp = sigma * 3
    stride = np.asarray(image_size) / np.asarray(heatmap_size)
    size = 2 * tmp + 1
    xs = np.arange(0, size, 1, np.float32)
    g = np.exp(-((xs - size // 2) ** 2 + (xs[:, None] - size // 2) ** 2) / (2 * sigma ** 2))
    for j in range(J):
        mu_x = int(joints[j][0] / stride[0] + 0.5)
        mu_y = int(joints[j][1] / stride[1] + 0.5)
        ul = [int(mu_x - tmp), int(mu_y - tmp)]
        br = [int(mu_x + tmp + 1), int(mu_y + tmp + 1)]
        if ul[0] >= W or ul[1] >= H or br[0] < 0 or br[1] < 0:
            tw[j] = 0
            continue
        gx = max(0, -ul[0]), min(br[0], W) - ul[0]
        gy = max(0, -ul[1]), min(br[1], H) - ul[1]
        ix = max(0, ul[0]), min(br[0], W)
        iy = max(0, ul[1]), min(br[1], H)
        if tw[j] > 0.5:
            target[j][iy[0]:iy[1], ix[0]:ix[1]] = g[gy[0]:gy[1], gx[0]:gx[1]]
    if joints_weight is not None:
        tw = np.multiply(tw, joints_weight)
    return target, tw
