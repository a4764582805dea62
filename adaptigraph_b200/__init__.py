"""adaptigraph_b200 — B200-native engine for AdaptiGraph's particle-graph dynamics hot path.

Importing the compute API loads libadaptigraph_b200.so and fails loudly if it is missing.
`adaptigraph_b200.synthetic` (workload generators) is importable without it.
"""
__all__ = ["DynamicsPredictor", "GraphedRollout", "EdgeList", "build_edges", "construct_edges_from_states",
           "construct_edges_from_states_batch", "edges_from_onehots", "pad_torch", "truncate_graph",
           "fps", "fps_rad_idx", "farthest_point_sampler", "fps_batch", "relation_lists", "collate_relation_lists",
           "DynDataset", "make_loader", "load_pairs", "load_dataset", "load_positions"]


def __getattr__(name):
    if name in ("DynamicsPredictor", "GraphedRollout"):
        from . import model
        return getattr(model, name)
    if name in ("EdgeList", "build_edges", "construct_edges_from_states", "construct_edges_from_states_batch",
                "edges_from_onehots", "relation_lists", "collate_relation_lists"):
        from . import graph
        return getattr(graph, name)
    if name in ("pad_torch", "truncate_graph"):
        from . import utils
        return getattr(utils, name)
    if name in ("fps", "fps_rad_idx", "farthest_point_sampler", "fps_batch"):
        from . import sampling
        return getattr(sampling, name)
    if name in ("DynDataset", "make_loader", "load_pairs", "load_dataset", "load_positions"):
        from . import dataset
        return getattr(dataset, name)
    raise AttributeError(name)
