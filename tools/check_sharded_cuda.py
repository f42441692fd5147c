#!/usr/bin/env python
"""Run under torchrun on >= 2 GPUs: the NCCL-sharded training set equals the single-GPU one.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29517 tools/check_sharded_cuda.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sbi_for_diffusion_models_b200 import data_simulator as ds  # noqa: E402
from sbi_for_diffusion_models_b200 import proposals  # noqa: E402
from sbi_for_diffusion_models_b200.sharding import simulate_training_set_sharded  # noqa: E402


class Prior:
    def sample(self, shape):
        return torch.rand((shape[0], 5)) * torch.tensor([1.0, 1.0, 2.0, 20.0, 1.0]) + torch.tensor([0.0, 0.0, 0.0, 5.0, 0.0])

    def log_prob(self, th):
        return torch.zeros(th.shape[:-1])


def make():
    torch.manual_seed(123)
    return proposals.ExtendedProposal(Prior(), proposals.PulseSequenceProposal(P=80, p_success=0.75, seed=4))


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, bs = 100003, 4096
    z, x = simulate_training_set_sharded(make(), n, bs, torch.device("cuda", local), mu_sensory=1.0, p_success=0.75,
                                         P=80, log_rt=False, seed=77)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        z1, x1 = ds.simulate_training_set_with_conditions(make(), n, bs, "cpu", mu_sensory=1.0, p_success=0.75, P=80,
                                                          log_rt=False, seed=77)
        ok = torch.equal(z.cpu(), z1) and torch.equal(x.cpu(), x1)
        print(f"sharded over {dist.get_world_size()} GPUs == single GPU: {ok}; outcomes "
              f"{torch.bincount(x1[:, 1].long(), minlength=3).tolist()}", flush=True)
    # SBC: datasets sharded over the ranks == the same run on one rank
    from sbi_for_diffusion_models_b200.mnle import run_sbc
    from sbi_for_diffusion_models_b200.mnle_net import DeviceMNLE, PackedMNLE
    from sbi_for_diffusion_models_b200.priors import build_prior_theta
    from sbi_for_diffusion_models_b200.run_config import RunConfig
    import bench
    est = DeviceMNLE(PackedMNLE.from_params(bench.random_mnle_params(0)))
    cfg = RunConfig(WARMUP_STEPS=3, NUM_TRIALS_OBS=10)
    solo = dist.new_group([0])
    out = run_sbc(cfg, prior_theta=build_prior_theta(), density_estimator=est, num_datasets=7,
                  posterior_samples_per_dataset=130, seed=5, save=False)
    if rank == 0:
        one = run_sbc(cfg, prior_theta=build_prior_theta(), density_estimator=est, num_datasets=7,
                      posterior_samples_per_dataset=130, seed=5, save=False, group=solo)
        same = (out["ranks"] == one["ranks"]).all() and all(torch.equal(a, b) for a, b in zip(out["all_samples"], one["all_samples"]))
        print(f"SBC sharded over {dist.get_world_size()} GPUs == single GPU: {bool(same)}", flush=True)
        ok = ok and bool(same)
    # potential: chains split over the ranks == all chains on one rank (configs[3] shape)
    from sbi_for_diffusion_models_b200.sharding import loglik_sum_sharded
    from sbi_for_diffusion_models_b200.data_simulator import simulate_observed_session
    x_o, pulses_o = simulate_observed_session(torch.tensor([0.45, 0.6, 1.3, 14.0, 0.25]), 50, "cuda", mu_sensory=1.0,
                                              p_success=0.75, P=80, seed=123, log_rt=False, noise_seed=11)
    torch.manual_seed(5)
    theta = build_prior_theta().sample((1027,)).cuda()
    pot = loglik_sum_sharded(lambda th: est.loglik_sum(th, x_o.cuda(), pulses_o.cuda()), theta)
    if rank == 0:
        same = torch.equal(pot, est.loglik_sum(theta, x_o.cuda(), pulses_o.cuda()))
        print(f"potential sharded over {dist.get_world_size()} GPUs == single GPU: {bool(same)}", flush=True)
        ok = ok and bool(same)
    # fused all-gather: the kernels store every result into all ranks' gathered array (peer memory over NVLink)
    from sbi_for_diffusion_models_b200.sharding import PeerGather
    from sbi_for_diffusion_models_b200.simulator import simulate_trials
    world = dist.get_world_size()
    m = 50000
    torch.manual_seed(9)
    z_full = make().sample((world * m,)).cuda()                 # same z on every rank
    pg = PeerGather(m)
    pg.x_all.fill_(-1.0)
    torch.cuda.synchronize()
    dist.barrier()
    mine = z_full[rank * m:(rank + 1) * m]
    simulate_trials(mine[:, :5], mine[:, 5:], seed=31, trial_offset=rank * m, out=pg.local, peer_blocks=pg.peers)
    pg.barrier()
    torch.cuda.synchronize()
    want = simulate_trials(z_full[:, :5], z_full[:, 5:], seed=31)
    same = torch.equal(pg.x_all, want)
    print(f"[rank {rank}] fused peer-store gather over {world} GPUs == single GPU: {bool(same)}", flush=True)
    ok = ok and bool(same)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
