"""Autograd bridge for DynamicsPredictor.forward (training unroll, train.py:90-112)."""


def forward_with_grad(model, state, attrs, action, p_instance, physics_param, edges):
    raise NotImplementedError(
        "adaptigraph_b200: the backward kernels for DynamicsPredictor.forward are not built yet; "
        "call the model under torch.no_grad() (forward / rollout / MPC planning are supported).")
