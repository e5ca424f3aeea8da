"""Drop-in for the reference's `src/lib/Estimators.py` (parameter estimators on the state-estimation coefficients).

`EstimatorLinear` / `EstimatorInv` (reference :24-37) run the batched contraction kernel `romhc_estimator`.
`EstimatorNear` is the argmax lookup of :18-21.  The sklearn-based `EstimatorTree` / `EstimatorNN` (:50-97) are
never constructed by the reference and are outside the accelerated path.
"""
import ctypes as C

import numpy as np


def _contract(c_values, a_values_base, invert):
    import torch
    from .. import _lib
    if not torch.cuda.is_available():
        raise _lib.RomhcError("no CUDA device: the ROMHighContrast B200 path has no CPU fallback")
    c = np.asarray(c_values, dtype=np.float64)
    a = np.asarray(a_values_base, dtype=np.float64)
    tail_shape = a.shape[1:]
    n = a.shape[0]
    c2 = np.ascontiguousarray(c.reshape(n, -1))
    dev = torch.device("cuda", torch.cuda.current_device())
    cd = torch.as_tensor(c2, device=dev)
    ad = torch.as_tensor(np.ascontiguousarray(a.reshape(n, -1)), device=dev)
    K, nb = cd.shape[1], ad.shape[1]
    out = torch.empty(K, nb, dtype=torch.float64, device=dev)
    _lib.call("romhc_estimator", C.c_void_p(cd.data_ptr()), K, n, C.c_void_p(ad.data_ptr()), nb, 1 if invert else 0,
              C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return out.cpu().numpy().reshape(c.shape[1:] + tail_shape)


class Estimator:
    def __init__(self, a_values_base):
        self.a_values_base = a_values_base

    def fit(self, c_values, a_values):
        return self

    def estimate_parameter(self, c_values):
        pass


class EstimatorNear(Estimator):
    def estimate_parameter(self, c_values):
        super(EstimatorNear, self).estimate_parameter(c_values)
        return self.a_values_base[np.argmax(c_values, axis=1), :]


class EstimatorLinear(Estimator):
    def estimate_parameter(self, c_values):
        """einsum("bi,b...->i...", c, a_basis)   (reference :24-27)"""
        super(EstimatorLinear, self).estimate_parameter(c_values)
        return _contract(c_values, self.a_values_base, invert=False)


class EstimatorInv(Estimator):
    def __init__(self, a_values_base):
        super().__init__(a_values_base)
        self.inv_a_values_base = 1.0 / np.array(self.a_values_base)

    def estimate_parameter(self, c_values):
        """1 / einsum("bi,b...->i...", c, 1 / a_basis)   (reference :30-37)"""
        super(EstimatorInv, self).estimate_parameter(c_values)
        return _contract(c_values, self.a_values_base, invert=True)


class EstimatorTree(Estimator):
    def __init__(self, a_values_base):
        raise Exception("Not implemented.")


class EstimatorNN(Estimator):
    def __init__(self, a_values_base, hidden_layer_sizes):
        raise Exception("Not implemented.")
