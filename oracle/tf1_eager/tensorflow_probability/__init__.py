"""tfp.math.pinv for the eager TensorFlow restatement next door (test infrastructure only)."""
import numpy as np
import torch

import tensorflow as tf


class math(object):
    @staticmethod
    def pinv(a, rcond=None):
        """Moore-Penrose inverse through the SVD; singular values <= rcond * largest are dropped, with
        rcond = 10 * max(rows, cols) * machine epsilon of the matrix dtype by default.  The reference hands it the float32
        mel matrix, so the decomposition runs in float32 (numpy's LAPACK gesdd, the routine TF's CPU Svd kernel wraps through
        Eigen's BDCSVD family; the cut below decides the rank, 726 of 1024 at the reference's size)."""
        t = tf._raw(a)
        if rcond is None:
            rcond = 10.0 * max(t.shape[-2:]) * float(np.finfo(np.float32 if t.dtype == torch.float32 else np.float64).eps)
        m = t.numpy()
        u, s, vt = np.linalg.svd(m, full_matrices=False)
        keep = s > rcond * s.max()
        inv = np.where(keep, 1.0 / np.where(keep, s, 1.0), 0.0).astype(m.dtype)
        return tf.Tensor(torch.from_numpy(((vt.T * inv) @ u.T).astype(m.dtype)))
