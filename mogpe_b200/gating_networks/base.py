"""mogpe/gating_networks/base.py."""
from ..gpflow_lite import Module


class GatingNetworkBase(Module):
    def predict_fs(self, Xnew, **kwargs):
        raise NotImplementedError

    def predict_mixing_probs(self, Xnew, **kwargs):
        raise NotImplementedError
