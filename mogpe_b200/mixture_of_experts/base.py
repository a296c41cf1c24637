"""mogpe/mixture_of_experts/base.py: MixtureOfExperts base (predict_y as MixtureSameFamily)."""
import numpy as np

from .._model_core import MixtureDistribution, Normal
from ..experts import ExpertsBase
from ..gating_networks import GatingNetworkBase
from ..gpflow_lite import Module


class MixtureOfExperts(Module):
    def __init__(self, gating_network: GatingNetworkBase, experts: ExpertsBase):
        assert isinstance(gating_network, GatingNetworkBase)
        self.gating_network = gating_network
        assert isinstance(experts, ExpertsBase)
        self.experts = experts
        self.num_experts = experts.num_experts

    def predict_mixing_probs(self, Xnew, **kwargs):
        """[N, K] (base.py:58-76; the reference evaluates the gating network twice -- once is enough)."""
        mixing_probs = self.gating_network.predict_mixing_probs(Xnew, **kwargs)
        assert mixing_probs.shape == (np.shape(Xnew)[0], self.num_experts), \
            "Mixing probabilities dimensions (from gating network) should be [num_data, num_experts]"
        return mixing_probs

    def predict_experts_dists(self, Xnew, **kwargs) -> Normal:
        return self.experts.predict_dists(Xnew, **kwargs)

    def predict_y(self, Xnew, **kwargs) -> MixtureDistribution:
        raise NotImplementedError

    def predict_y_samples(self, Xnew, num_samples: int = 1, **kwargs):
        return self.predict_y(Xnew, **kwargs).sample(num_samples)
