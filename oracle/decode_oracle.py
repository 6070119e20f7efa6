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
