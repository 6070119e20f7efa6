"""Hot-path subset of the reference's lib/utils/transforms.py: flip_back (transforms.py:15-29). On CUDA tensors
it is a pure index permutation done on the device (no numpy round-trip)."""
import numpy as np
import torch


def flip_back(output_flipped, matched_parts):
    if isinstance(output_flipped, np.ndarray):
        assert output_flipped.ndim == 4, 'output_flipped should be [batch_size, num_joints, height, width]'
        t = torch.from_numpy(np.ascontiguousarray(output_flipped))
        return flip_back(t, matched_parts).numpy()
    J = output_flipped.shape[1]
    perm = list(range(J))
    for a, b in matched_parts:
        perm[a], perm[b] = b, a
    idx = torch.tensor(perm, device=output_flipped.device)
    return output_flipped.flip(3).index_select(1, idx)
