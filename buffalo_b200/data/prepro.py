"""Value pre-processors applied to the stored values (buffalo/data/prepro.py): same class names/options."""
import numpy as np


class PreProcess(object):
    def __init__(self, opt):
        self.opt = opt

    def pre(self, header):
        pass

    def __call__(self, v):
        return v

    def post(self, db):
        pass


class OneBased(PreProcess):
    def __call__(self, v):
        v[:] = 1.0
        return v


class MinMaxScalar(PreProcess):
    def __init__(self, opt):
        super().__init__(opt)
        self.value_min, self.value_max = float("inf"), 0.0

    def __call__(self, V):
        if len(V):
            self.value_min = min(self.value_min, float(np.min(V)))
            self.value_max = max(self.value_max, float(np.max(V)))
        return V

    def post(self, db):
        span = self.value_max - self.value_min
        if span < 1e-8:
            db["val"][:] = 1.0 * self.opt.max
            return
        v = db["val"][:]
        db["val"][:] = (v - self.value_min) / span * (self.opt.max - self.opt.min) + self.opt.min


class ImplicitALS(PreProcess):
    def __call__(self, V):
        return np.log(1 + V / self.opt.epsilon)


class SPPMI(PreProcess):
    pass
