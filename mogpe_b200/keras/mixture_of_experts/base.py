"""mogpe/keras/mixture_of_experts/base.py: MixtureOfExpertsBase (predict_y as a mixture of the experts)."""
from typing import List

from ..._model_core import MixtureDistribution
from ...gpflow_lite import Module
from ..experts import ExpertBase
from ..gating_networks import GatingNetworkBase


class MixtureOfExpertsBase(Module):
    def __init__(self, experts_list: List[ExpertBase], gating_network: GatingNetworkBase, name: str = "MixtureOfExperts"):
        for expert in experts_list:
            assert isinstance(expert, ExpertBase)
        assert isinstance(gating_network, GatingNetworkBase)
        self._experts_list = list(experts_list)
        self._gating_network = gating_network
        assert gating_network.num_experts == self.num_experts
        self.name = name

    @property
    def experts_list(self):
        return self._experts_list

    @property
    def gating_network(self):
        return self._gating_network

    @property
    def num_experts(self) -> int:
        return len(self._experts_list)

    def predict_experts_dists(self, Xnew, **kwargs):
        return [expert(Xnew, **kwargs) for expert in self.experts_list]

    def predict_y(self, Xnew, **kwargs) -> MixtureDistribution:
        raise NotImplementedError
