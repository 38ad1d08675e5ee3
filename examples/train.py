#!/usr/bin/env python
"""Actor / replay / learner pipeline on the B200 learner, with the process topology of the reference's train.py
(train.py:20-44 upstream: shared-memory model, N actor processes, one buffer process, learner in the main process).

The reference's own train.py runs unmodified with `PYTHONPATH=<repo>/dropin:<repo>`; this script is
the same flow with command-line overrides for the config constants, usable where gym/ALE are absent (synthetic env).
"""
import argparse
import os
import random
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from r2d2_b200 import config  # noqa: E402


def epsilon_for(actor_id: int, num_actors: int, base_eps: float = config.base_eps, alpha: float = config.alpha) -> float:
    """Per-actor exploration rate base_eps^(1 + i/(N-1)*alpha) (train.py:15-17); a single actor gets base_eps."""
    return base_eps ** (1 + (actor_id / (num_actors - 1) if num_actors > 1 else 0) * alpha)


def _gpu_actor_main(epsilons, model, sample_queue, obs_shape, block_length, device_index, overrides):
    """child process (spawned: it needs its own CUDA context): all environments behind one batched GPU inference"""
    for k, v in overrides.items():
        setattr(config, k, v)
    torch.set_num_threads(1)
    from r2d2_b200.worker import VectorActor
    VectorActor(epsilons, model, sample_queue, obs_shape=obs_shape, block_length=block_length,
                device=torch.device("cuda", device_index)).run()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--actors", type=int, default=config.num_actors)
    ap.add_argument("--gpu-actors", action="store_true",
                    help="step all actors' environments in ONE process with a batched GPU inference per step "
                         "(worker.VectorActor) instead of one CPU-inference process per actor")
    ap.add_argument("--synthetic-env", action="store_true",
                    help="no Atari emulator on this box: run on r2d2_b200.environment.SyntheticAtariEnv (sets R2D2_SYNTHETIC_ENV=1 "
                         "for this process and the actor processes); without it a missing gym/ALE is an error, as upstream")
    for name in ("training_steps", "learning_starts", "buffer_capacity", "batch_size", "log_interval", "block_length",
                 "burn_in_steps", "learning_steps", "forward_steps", "save_interval"):
        ap.add_argument("--" + name.replace("_", "-"), type=int, default=None)
    args = ap.parse_args()
    if args.synthetic_env:
        os.environ["R2D2_SYNTHETIC_ENV"] = "1"
    overrides = {name: val for name, val in vars(args).items() if name not in ("actors", "gpu_actors", "synthetic_env") and val is not None}
    for name, val in overrides.items():
        setattr(config, name, val)

    torch.manual_seed(0)
    np.random.seed(0)
    random.seed(0)
    torch.set_num_threads(1)

    from r2d2_b200.environment import create_env
    from r2d2_b200.model import Network
    from r2d2_b200.worker import Actor, Learner, ReplayBuffer

    model = Network(create_env().action_space.n)
    model.share_memory()
    spawn = mp.get_context("spawn")
    sample_queues = [spawn.Queue()] if args.gpu_actors else [mp.Queue() for _ in range(args.actors)]
    batch_queue, priority_queue = mp.Queue(8), mp.Queue(8)

    buffer = ReplayBuffer(sample_queues, batch_queue, priority_queue, buffer_capacity=config.buffer_capacity,
                          batch_size=config.batch_size)
    learner = Learner(batch_queue, priority_queue, model, save_interval=config.save_interval)     # CUDA is initialised here
    if args.gpu_actors:
        eps = [epsilon_for(i, args.actors) for i in range(args.actors)]
        procs = [spawn.Process(target=_gpu_actor_main, daemon=True,
                               args=(eps, model, sample_queues[0], tuple(config.obs_shape), config.block_length,
                                     torch.cuda.current_device(), overrides))]
    else:
        actors = [Actor(epsilon_for(i, args.actors), model, sample_queues[i], block_length=config.block_length)
                  for i in range(args.actors)]
        procs = [mp.Process(target=a.run, daemon=True) for a in actors]
    for p in procs:
        p.start()
    buffer_proc = mp.Process(target=buffer.run)
    buffer_proc.start()

    learner.run()

    buffer_proc.join(timeout=3 * config.log_interval + 5)
    if buffer_proc.is_alive():
        buffer_proc.terminate()
    for p in procs:
        p.terminate()
    print(f"done: {learner.num_updates} updates, replay size {len(learner.replay)}")


if __name__ == "__main__":
    main()
