"""Drop-in for the box-NMS entry points of the reference's lib/nms (gpu_nms.pyx:19-34 `gpu_nms`, nms.py:27-32
`gpu_nms_wrapper`), backed by libfpd_b200's sm_100a kernels (csrc/nms.cu). OKS-NMS (nms.py:75-124) stays host
numpy in the reference and is outside this path."""
import ctypes

import numpy as np


def gpu_nms(dets, thresh, device_id=0):
    """dets: float32 [N,5] (x1,y1,x2,y2,score). Returns the kept indices into `dets`, highest score first --
    same contract as the Cython wrapper: sort by score, run `_nms` on the sorted boxes, map back through `order`."""
    from fpd_b200 import _native as N
    dets = np.ascontiguousarray(dets, dtype=np.float32)
    n, d = dets.shape
    if n == 0:
        return []
    order = dets[:, 4].argsort()[::-1].astype(np.int32)
    sorted_dets = np.ascontiguousarray(dets[order, :])
    keep = np.zeros(n, dtype=np.int32)
    num = ctypes.c_int(0)
    N.check(N.lib().fpd_nms_host(keep.ctypes.data_as(ctypes.c_void_p), ctypes.cast(ctypes.byref(num), ctypes.c_void_p),
                                 sorted_dets.ctypes.data_as(ctypes.c_void_p), n, d, float(thresh), int(device_id)),
            "nms_host")
    return list(order[keep[:num.value]])


def gpu_nms_wrapper(thresh, device_id):
    def _nms(dets):
        return gpu_nms(dets, thresh, device_id)
    return _nms
