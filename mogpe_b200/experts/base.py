"""mogpe/experts/base.py: abstract expert(s) interfaces."""
from typing import List

from ..gpflow_lite import Module


class ExpertBase(Module):
    def predict_dist(self, Xnew, **kwargs):
        raise NotImplementedError


class ExpertsBase(Module):
    """Container of experts (mogpe/experts/base.py:35-79)."""

    def __init__(self, experts_list: List[ExpertBase] = None, name="Experts"):
        assert isinstance(experts_list, list), "experts_list should be a list of ExpertBase instances"
        for expert in experts_list:
            assert isinstance(expert, ExpertBase)
        self.num_experts = len(experts_list)
        self.experts_list = experts_list
        self.name = name

    def predict_dists(self, Xnew, **kwargs):
        raise NotImplementedError
