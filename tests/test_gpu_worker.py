"""The public worker / model surface on the GPU: Learner.update_from_batch on reference-format 14-tuples,
Network.calculate_q_/calculate_q, and the HBM-replay learner loop."""
import os
import queue

import numpy as np
import pytest
import torch

from oracle import synth
from oracle.learner import init_params
from helpers import build_oracle_replay, sample_with_seed, A

pytestmark = pytest.mark.gpu


def _tuple14(d):
    t = torch.from_numpy
    return (t(d["obs"]), t(d["last_action"]), t(d["last_reward"]), t(np.ascontiguousarray(d["hidden"])).transpose(0, 1),
            t(d["action"]).unsqueeze(1), t(d["n_step_reward"]), t(d["gamma"]), t(d["burn_in"]), t(d["learning"]), t(d["forward"]),
            d["idxes"], t(d["is_weights"]), d["old_ptr"], np.int32(d["env_steps"]))


def _network(params):
    from r2d2_b200.model import Network
    net = Network(A)
    net.load_state_dict(params)
    return net


def test_learner_api_on_reference_tuples_matches_golden(golden_dir, monkeypatch):
    from r2d2_b200 import config
    from r2d2_b200.worker import Learner
    g = np.load(os.path.join(golden_dir, "learner_ragged.npz"))
    batch_size, K, bl, ls, bi, fs, seed0, num_blocks = (int(x) for x in g["meta"])
    monkeypatch.setattr(config, "batch_size", batch_size)
    rb, _ = build_oracle_replay(synth.RAGGED_SCRIPT, num_blocks, batch_size)
    model = _network(init_params(A, seed=3))
    pq = queue.Queue()
    learner = Learner(queue.Queue(), pq, model)
    learner._start_time = 0.0
    for k in range(K):
        idxes, prio, old_ptr, loss = learner.update_from_batch(_tuple14(sample_with_seed(rb, seed0 + k)))
        np.testing.assert_array_equal(idxes, g[f"k{k}_out_idxes"])
        np.testing.assert_allclose(prio, g[f"k{k}_out_priorities"], atol=1e-4, rtol=0)        # north-star bar
        assert abs(loss - float(g[f"k{k}_out_loss"])) < 2e-5
        assert prio.dtype == np.float32 and isinstance(loss, float)
    assert learner.num_updates == K
    sd = learner.state_dict()
    for n in sd:
        np.testing.assert_allclose(sd[n].flatten()[:16].cpu().numpy(), g[f"k{K-1}_phead_{n}"], atol=2e-5, rtol=0)


def test_network_calculate_q_surface(golden_dir):
    g = np.load(os.path.join(golden_dir, "learner_ragged.npz"))
    batch_size, K, bl, ls, bi, fs, seed0, num_blocks = (int(x) for x in g["meta"])
    rb, _ = build_oracle_replay(synth.RAGGED_SCRIPT, num_blocks, batch_size)
    d = sample_with_seed(rb, seed0)
    net = _network(init_params(A, seed=3)).cuda()
    t = lambda x: torch.from_numpy(x).cuda()
    obs = t(d["obs"]).float() / 255                                   # what worker.py:336,342 hands to the network
    hid = t(np.ascontiguousarray(d["hidden"])).transpose(0, 1)
    hidden_state = (hid[:1], hid[1:])
    args = (obs, t(d["last_action"]).float(), t(d["last_reward"]), hidden_state, t(d["burn_in"]), t(d["learning"]))
    q_shift = net.calculate_q_(*args, t(d["forward"]))
    q = net.calculate_q(*args)
    np.testing.assert_allclose(q_shift.cpu().numpy(), g["k0_out_qn_online"], atol=1e-4, rtol=0)
    np.testing.assert_allclose(q.cpu().numpy(), g["k0_out_q"], atol=1e-4, rtol=0)


def test_learner_loop_on_hbm_replay(monkeypatch):
    """Blocks in -> HBM replay -> sample/update/priority-update entirely on the device, through Learner.run."""
    from r2d2_b200 import config
    from r2d2_b200.worker import BLOCK_MSG, STATS_MSG, Learner, LocalBuffer
    monkeypatch.setattr(config, "batch_size", 8)
    monkeypatch.setattr(config, "buffer_capacity", 8 * 400)
    monkeypatch.setattr(config, "learning_starts", 400)
    monkeypatch.setattr(config, "training_steps", 6)
    bq, pq = queue.Queue(), queue.Queue()
    model = _network(init_params(A, seed=5))
    model.share_memory()
    before = {k: v.clone() for k, v in model.state_dict().items()}
    learner = Learner(bq, pq, model, save_interval=10 ** 9)
    for seed, steps, done in synth.RAGGED_SCRIPT:
        lb = LocalBuffer(A)
        for blk, prio, ep in synth.drive_actor(lb, seed, steps, done, A):
            bq.put((BLOCK_MSG, blk, prio, ep))
    learner.run()
    assert learner.num_updates == 6 and len(learner.replay) > 400
    msg = None
    while not pq.empty():
        msg = pq.get()
    assert msg[0] == STATS_MSG and msg[1] == 6 and np.isfinite(msg[2])
    changed = sum(float((model.state_dict()[k] - before[k]).abs().sum()) for k in before)
    assert changed > 0                                               # weights were published to the shared model (every 4 updates)
    leaves = learner.replay.tree.ptree[learner.replay.tree.num_nodes // 2:]
    assert np.isfinite(leaves).all() and (leaves >= 0).all()


def test_learner_run_consumes_reference_tuples(golden_dir, monkeypatch):
    """Learner.run() fed reference-format 14-tuples through its batch queue (the topology where a CPU-side sampler
    produces host batches): priorities come back on the priority queue in the reference's (idxes, priorities, old_ptr,
    loss) format and match the recorded reference outputs."""
    from r2d2_b200 import config
    from r2d2_b200.worker import Learner
    g = np.load(os.path.join(golden_dir, "learner_ragged.npz"))
    batch_size, K, bl, ls, bi, fs, seed0, num_blocks = (int(x) for x in g["meta"])
    monkeypatch.setattr(config, "batch_size", batch_size)
    monkeypatch.setattr(config, "training_steps", K)
    rb, _ = build_oracle_replay(synth.RAGGED_SCRIPT, num_blocks, batch_size)
    bq, pq = queue.Queue(), queue.Queue()
    for k in range(K):
        bq.put(_tuple14(sample_with_seed(rb, seed0 + k)))
    learner = Learner(bq, pq, _network(init_params(A, seed=3)), save_interval=10 ** 9)
    learner.run()
    for k in range(K):
        idxes, prio, old_ptr, loss = pq.get(timeout=5)
        np.testing.assert_array_equal(idxes, g[f"k{k}_out_idxes"])
        np.testing.assert_allclose(prio, g[f"k{k}_out_priorities"], atol=1e-4, rtol=0)
        assert abs(loss - float(g[f"k{k}_out_loss"])) < 2e-5


def test_async_weight_publication_reaches_the_shared_model():
    """store_weights (worker.py:306-307): the shared CPU model the actors read receives the online parameters, via the
    side-stream D2H + daemon thread, without the learner waiting on it."""
    import torch
    from r2d2_b200 import config
    from r2d2_b200.model import Network
    from r2d2_b200.worker import Learner
    from oracle.learner import init_params
    A, C = 6, 1
    config.obs_shape = (C, 84, 84)
    shared = Network(A, obs_shape=(C, 84, 84))
    shared.load_state_dict(init_params(A, in_channels=C, seed=3))
    learner = Learner(None, None, shared, save_interval=10 ** 9, device=torch.device("cuda", 0))
    with torch.no_grad():
        learner.core.online.flat.mul_(1.5).add_(0.25)                  # pretend an update happened
    want = {k: v.clone().cpu() for k, v in learner.core.online.views.items()}
    learner.store_weights(wait=True)
    assert learner._publisher.published == 1
    got = shared.state_dict()
    for k, v in want.items():
        assert torch.equal(got[k].cpu(), v), k
    assert learner._publisher.publish() is True                         # idle again
    learner._publisher.wait()


def test_batched_policy_matches_cpu_network_forward():
    """N actors stepped by ONE GPU forward (actor_inference.BatchedPolicy) == N batch-1 CPU Network.forward calls
    (model.py:65-79), over several steps with the recurrent state carried on the device."""
    import numpy as np
    import torch
    from r2d2_b200 import config
    from r2d2_b200.actor_inference import BatchedPolicy
    from r2d2_b200.model import AgentState, Network
    from oracle.learner import init_params
    for C, N, A in [(1, 5, 6), (4, 16, 9)]:
        config.obs_shape = (C, 84, 84)
        net = Network(A, obs_shape=(C, 84, 84))
        net.load_state_dict(init_params(A, in_channels=C, seed=11))
        net.eval()
        pol = BatchedPolicy(A, N, obs_shape=(C, 84, 84), device=torch.device("cuda", 0))
        pol.load_state_dict(net.state_dict())
        rng = np.random.default_rng(5)
        states = [AgentState(torch.zeros(1, C, 84, 84), A) for _ in range(N)]
        hidden = torch.zeros(N, 2, 512)
        last_action = np.zeros(N, dtype=np.int64)
        last_reward = np.zeros(N, dtype=np.float32)
        first = True
        for step in range(4):
            obs = rng.integers(0, 256, size=(N, C, 84, 84), dtype=np.uint8)
            la = np.zeros((N, A), dtype=np.uint8)
            if not first:
                la[np.arange(N), last_action] = 1
            q_gpu, h_gpu = pol.step(obs, la, last_reward, hidden if first else None)
            q_gpu, h_gpu = q_gpu.cpu(), h_gpu.cpu()
            for i in range(N):
                st = states[i]
                st.obs = torch.from_numpy(obs[i]).unsqueeze(0).float()
                st.last_action = torch.from_numpy(la[i]).float().unsqueeze(0)
                st.last_reward = torch.tensor([[last_reward[i]]])
                with torch.no_grad():
                    q_cpu, (h, c) = net(st)
                st.hidden_state = (h, c)
                assert torch.allclose(q_gpu[i], q_cpu[0], atol=2e-5, rtol=1e-5), (C, step, i, (q_gpu[i] - q_cpu[0]).abs().max())
                assert torch.allclose(h_gpu[i, 0], h.reshape(-1), atol=2e-5) and torch.allclose(h_gpu[i, 1], c.reshape(-1), atol=2e-5)
            last_action = q_gpu.argmax(1).numpy()
            last_reward = rng.normal(size=N).astype(np.float32)
            first = False


def test_vector_actor_ships_the_same_blocks_as_cpu_actors():
    """worker.VectorActor (one batched GPU inference per step for N envs) vs N reference-style `Actor`s (batch-1 CPU
    inference) on action-independent synthetic environments with epsilon 0: identical frames/rewards/segmentation,
    hidden states and initial priorities within the inference tolerance, including blocks cut mid-episode
    (worker.py:548-552) whose bootstrap Q comes from the next batched step."""
    import queue
    import numpy as np
    import torch
    from r2d2_b200 import config
    from r2d2_b200.environment import SyntheticAtariEnv
    from r2d2_b200.model import Network
    from r2d2_b200.worker import Actor, LocalBuffer, VectorActor
    from oracle.learner import init_params
    A, C, N, STEPS, BL = 9, 1, 3, 230, 80
    config.obs_shape = (C, 84, 84)
    shared = Network(A, obs_shape=(C, 84, 84))
    shared.load_state_dict(init_params(A, in_channels=C, seed=21))
    shared.eval()
    mk = lambda i: SyntheticAtariEnv(A, (C, 84, 84), mean_episode_len=150, seed=300 + i)

    shipped = [[] for _ in range(N)]
    va = VectorActor([0.0] * N, shared, queue.Queue(), obs_shape=(C, 84, 84), max_episode_steps=10 ** 6, block_length=BL,
                     device=torch.device("cuda", 0), envs=[mk(i) for i in range(N)])
    va._ship = lambda i, triple, over: shipped[i].append(triple)
    for _ in range(STEPS):
        va.step()

    total_blocks = 0
    for i in range(N):
        q = queue.Queue()
        ac = Actor(0.0, shared, q, obs_shape=(C, 84, 84), max_episode_steps=10 ** 6, block_length=BL)
        ac.env = mk(i)
        ac.model.load_state_dict(shared.state_dict())
        ac.local_buffer = LocalBuffer(A, block_length=BL)
        ref = []
        ac._ship = lambda triple, over, ref=ref: ref.append(triple)
        while ac.actor_steps < STEPS + 200:
            ac.play_episode()
        assert len(shipped[i]) >= 2
        for (blk, prio, ret), (rblk, rprio, rret) in zip(shipped[i], ref):
            total_blocks += 1
            assert np.array_equal(blk.obs, rblk.obs) and np.array_equal(blk.last_reward, rblk.last_reward)
            assert blk.num_sequences == rblk.num_sequences
            for f in ("burn_in_steps", "learning_steps", "forward_steps", "n_step_reward", "gamma"):
                assert np.array_equal(getattr(blk, f), getattr(rblk, f)), f
            assert (blk.action == rblk.action).mean() >= 0.98            # argmax of Q: ties aside identical
            if np.array_equal(blk.action, rblk.action):
                assert np.array_equal(blk.last_action, rblk.last_action)
                assert np.allclose(blk.hidden, rblk.hidden, atol=1e-4)
                assert np.allclose(prio, rprio, atol=1e-4)
            assert (ret is None) == (rret is None)
    assert total_blocks >= 2 * N
