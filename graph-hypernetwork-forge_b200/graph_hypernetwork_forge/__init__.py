"""graph_hypernetwork_forge - B200-native build of the HyperGNN forward path.

Import-compatible with the reference package of the same name
(`from graph_hypernetwork_forge import WeightGenerator, HyperGNN, ToyKnowledgeGraph`);
the forward pass runs in hand-written sm_100a kernels (libghf_b200.so, C ABI in
include/ghf_b200.h).  CUDA tensors only; forward only.
"""

__version__ = "0.2.0+b200.1"

from .models.weight_generator import WeightGenerator
from .models.hypergnn import HyperGNN
from .data.knowledge_graph import ToyKnowledgeGraph

__all__ = ["WeightGenerator", "HyperGNN", "ToyKnowledgeGraph"]
