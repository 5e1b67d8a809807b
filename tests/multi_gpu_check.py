"""Batch-sharded run == un-sharded run. Launch with torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/multi_gpu_check.py

Every rank runs the SAME sampler job on its batch slice of a global batch (global scale_noise
statistics exchanged through the NVLink peer mailboxes, or NCCL with SONAR_B200_NO_PEER=1); the
gathered result must equal the single-GPU run of the whole batch.
"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import sonar_b200 as sb  # noqa: E402


def main() -> None:
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    batch = 2 * world + 1  # ragged split on purpose
    torch.manual_seed(0)
    sigmas = torch.cat((torch.linspace(14.6, 0.5, 6), torch.zeros(1))).to(dev)
    x_image = (torch.randn(batch, 4, 32, 48) * 14.6).to(dev)

    def model(x, sigma, **_kw):
        return x * 0.9

    def run(sampler, x, **kw):
        torch.manual_seed(1234)  # replicated generators: every rank reserves the same global draws
        return sampler(model, x, sigmas, extra_args={"seed": 0}, disable=True, **kw)

    ng, sn = sb.noise_graph, sb.spectral_noise
    chain = ng.CustomNoiseChain()
    chain.add(sn.PowerNoiseItem(1.0, time_brownian=False, alpha=1.0, max_freq=0.7071, min_freq=0.0, stretch=1.0, rotate=0.0,
                                pnorm=2.0, mix=1.0, common_mode=0.0, channel_correlation="1, 1, 1, 1, 1, 1"))
    chain.add(ng.CustomNoiseItem(0.5, noise_type="pyramid"))
    jobs = {
        "euler_ancestral/fused gaussian": (sb.samplers.SonarEulerAncestral.sampler, {}),
        "dpmpp_sde/fused gaussian": (sb.samplers.SonarDPMPPSDE.sampler, {"sonar_params": {"noise_type": "gaussian"}}),
        "euler_ancestral/power+pyramid chain": (sb.samplers.SonarEulerAncestral.sampler, {"sonar_params": {"custom_noise": chain}}),
    }
    # north-star config 5, scaled down: frames_to_channels power noise as the custom noise of sonar_dpmpp_sde (5-D latent)
    video = ng.CustomNoiseChain()
    inner = ng.CustomNoiseChain()
    inner.add(sn.PowerNoiseItem(1.0, time_brownian=False, alpha=1.0, max_freq=0.7071, min_freq=0.0, stretch=1.0, rotate=0.0,
                                pnorm=2.0, mix=1.0, common_mode=0.0, channel_correlation="1, 1, 1, 1, 1, 1"))
    video.add(ng.CustomNoiseParametersNoise(
        1.0, noise=inner, normalize=None, override_device=None, override_dtype=None, frames_to_channels=True,
        ensure_square_aspect_ratio=False, fix_invalid=False, rng_mode="default", rng_offset_mode="disabled", rng_state_offset=0))
    x_video = (torch.randn(batch, 4, 3, 18, 20) * 14.6).to(dev)
    jobs["dpmpp_sde/C5 video power noise"] = (sb.samplers.SonarDPMPPSDE.sampler, {"sonar_params": {"custom_noise": video}})
    for name, (sampler, kw) in jobs.items():
        x_full = x_video if "C5" in name else x_image
        want = run(sampler, x_full, **kw)
        with sb.parallel.sharded(batch) as ctx:
            assert sb.parallel.device_barrier() == (ctx.peers is not None)  # device-side rendezvous, no host sync
            mine = run(sampler, sb.parallel.shard(x_full), **kw)
            got = sb.parallel.gather(mine)
            transport = "peer mailboxes" if ctx.peers is not None else "nccl"
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5, msg=lambda m: f"{name}: {m}")
        if rank == 0:
            print(f"OK {name}: {world} ranks ({transport}), max |diff| = {(got - want).abs().max().item():.3g}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
