"""One MPPI sampling iteration at the reference's planning configuration (config/planning/rope.yaml:31-46: max_nobj 200, max_nR 2000,
n_look_ahead 1, n_sample 20000 evaluated in chunks of n_sample_chunk = 500 by the reference): decode the sampled pushes, roll every
sample forward with per-step re-graphing (planning.dynamics = forward_dynamics.py:11-205), score with running_cost (plan.py:27-59:
chamfer to the target + collision and box penalties).  The engine evaluates the samples in chunks of --chunk (default 5000); the CPU
column times the oracle's restatement of the same drivers on a bounded sample of pushes with all host threads.  One JSON line."""
import argparse
import json
import os
import sys
import time
import types
from functools import partial

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import adaptigraph_b200 as agx  # noqa: E402
from adaptigraph_b200 import planning, rewards, synthetic as syn  # noqa: E402
from oracle import planning_oracle as po  # noqa: E402   (cpu baseline leg only)

ap = argparse.ArgumentParser()
ap.add_argument("--samples", type=int, default=20000)
ap.add_argument("--chunk", type=int, default=5000)
ap.add_argument("--particles", type=int, default=100)
ap.add_argument("--max-repeat", type=int, default=8)
ap.add_argument("--cpu-samples", type=int, default=64)
args = ap.parse_args()

material = "rope"
thr, topk, cta, _ = syn.MATERIALS[material]
pusher = [[0.0, 0.0]]
tc = dict(max_n=1, max_nR=2000, n_his=4, sim_real_ratio=10.0, push_length=0.1, pusher_points=pusher, gripper_enable=False, topk=topk,
          connect_tools_all=cta)
ppm = types.SimpleNamespace(task_config=tc, eef_num=1, material=material, material_dims={material: 1}, material_indices={material: 0},
                            physics_param={material: torch.tensor([0.4])}, adj_thresh=thr)
torch.manual_seed(0)
model = agx.DynamicsPredictor(*syn.configs(material, 3), "cuda").cuda().eval()
w = syn.make_workload(material, args.particles, 1, seed=2)
state = w.state[0, -1, :args.particles].contiguous()                        # (n_obj, 3) current object particles
target = (state + torch.tensor([0.3, 0.0, 0.2]))[None]                      # (1, n_obj, 3) goal configuration
bbox = torch.tensor([[float(state[:, 0].min()) - 1.0, float(state[:, 0].max()) + 1.0],
                     [float(state[:, 2].min()) - 1.0, float(state[:, 2].max()) + 1.0]])
g = torch.Generator().manual_seed(1)
n = args.samples
action = torch.stack([state[torch.randint(0, args.particles, (n,), generator=g), 0] + 0.2 * torch.randn(n, generator=g),
                      state[torch.randint(0, args.particles, (n,), generator=g), 2] + 0.2 * torch.randn(n, generator=g),
                      6.2832 * torch.rand(n, generator=g),
                      torch.randint(1, args.max_repeat + 1, (n,), generator=g).float()], -1)[:, None]      # (n, 1, 4)


def iteration():
    rew = []
    state_d, target_d, bbox_d = state.cuda(), target.cuda(), bbox.cuda()
    for c0 in range(0, n, args.chunk):
        a = action[c0:c0 + args.chunk].cuda(non_blocking=True)
        out = planning.dynamics(state_d, a, model, "cuda", ppm)
        r = rewards.running_cost(out["state_seqs"], out["action_seqs"], state_d, partial(rewards.chamfer, y=target_d), rewards.rope_penalty, bbox_d)
        rew.append(r["reward_seqs"])
    rew = torch.cat(rew)
    best = int(rew.argmax().item())                                          # the one host read an MPPI update needs
    return rew, best


for _ in range(2):
    iteration()
torch.cuda.synchronize()
t0 = time.perf_counter()
K = 3
for _ in range(K):
    rew, best = iteration()
torch.cuda.synchronize()
gpu_s = (time.perf_counter() - t0) / K
steps_total = float(action[:, 0, 3].sum()) * args.particles                   # particle-steps actually rolled (sum of repeat counts)

# CPU: the oracle's drivers on a bounded sample
torch.set_num_threads(os.cpu_count())
params = {k: v.detach().cpu() for k, v in model.state_dict().items()}
cfg = dict(pusher=pusher, ratio=10.0, push_length=0.1, gripper=False, thr=thr, topk=topk, cta=cta, n_his=4, phys=0.4)
a_cpu = action[:args.cpu_samples]
po.dynamics(params, 3, state, a_cpu[:8], cfg)
t0 = time.perf_counter()
ref, dec_cpu = po.dynamics(params, 3, state, a_cpu, cfg)
r_cpu = po.running_cost(ref, dec_cpu, state, partial(po.chamfer, y=target),
                        po.rope_penalty, bbox)
cpu_s = time.perf_counter() - t0
out = planning.dynamics(state.cuda(), a_cpu.cuda(), model, "cuda", ppm)
err = float((out["state_seqs"].cpu() - ref).abs().max())
print(json.dumps({
    "workload": f"rope {args.particles} particles, {n} sampled pushes (repeat 1..{args.max_repeat}), 1 look-ahead, chunks of {args.chunk}",
    "gpu_s_per_mppi_iteration": round(gpu_s, 4), "gpu_samples_per_s": round(n / gpu_s), "gpu_particle_steps_per_s": round(steps_total / gpu_s),
    "cpu_s_for_sample": round(cpu_s, 3), "cpu_samples": args.cpu_samples, "cpu_samples_per_s": round(args.cpu_samples / cpu_s, 1),
    "cpu_cores": os.cpu_count(), "speedup_samples_per_s": round((n / gpu_s) / (args.cpu_samples / cpu_s), 1),
    "max_abs_diff_vs_cpu_on_sample": err, "best_sample": best, "best_reward": float(rew[best])}))
