"""ToyKnowledgeGraph - the 8-node fixture of the reference (data/knowledge_graph.py:41-105).

Fixture data only (BASELINE config 1); nothing here is accelerated.  Tensors are created on the
CPU exactly like the reference (features: randn from a Generator seeded with 42); move them to
CUDA before calling the model.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

import torch

_NODES = ("Alice", "Bob", "Carol", "Dave", "Eve", "Acme Corp", "London", "Python")
# (head, relation, tail) by name; order defines the edge order and therefore relation ids
_TRIPLES = (
    ("Alice", "is spouse of", "Bob"), ("Bob", "is spouse of", "Alice"), ("Alice", "knows", "Carol"),
    ("Bob", "works with", "Dave"), ("Carol", "knows", "Dave"), ("Dave", "works at", "Acme Corp"),
    ("Alice", "works at", "Acme Corp"), ("Acme Corp", "located in", "London"),
    ("Alice", "has skill", "Python"), ("Dave", "has skill", "Python"), ("Carol", "is parent of", "Eve"),
)


def _default_edges() -> List[Tuple[int, int, str]]:
    pos = {name: i for i, name in enumerate(_NODES)}
    return [(pos[h], pos[t], rel) for h, rel, t in _TRIPLES]


@dataclass
class ToyKnowledgeGraph:
    feat_dim: int = 16
    node_names: List[str] = field(default_factory=lambda: list(_NODES))
    edge_data: List[tuple] = field(default_factory=_default_edges)

    def __post_init__(self) -> None:
        rng = torch.Generator()
        rng.manual_seed(42)
        self.node_features: torch.Tensor = torch.randn(len(self.node_names), self.feat_dim, generator=rng)
        heads, tails, rels = zip(*self.edge_data) if self.edge_data else ((), (), ())
        self.edge_index: torch.Tensor = torch.tensor([list(heads), list(tails)], dtype=torch.long)
        self.edge_texts: List[str] = list(rels)

    @property
    def num_nodes(self) -> int:
        return len(self.node_names)

    @property
    def num_edges(self) -> int:
        return self.edge_index.size(1)

    @property
    def relation_types(self) -> List[str]:
        return list(dict.fromkeys(self.edge_texts))

    def __repr__(self) -> str:
        return (f"ToyKnowledgeGraph(nodes={self.num_nodes}, edges={self.num_edges}, "
                f"relation_types={len(self.relation_types)})")
