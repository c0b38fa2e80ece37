"""Multi-GPU observation path, through the C ABI:
the newest-frame output of k_raster, rendering into a slice of a larger tensor,
k_stack_push against a plain-torch statement of FlattenFrameStack, the idempotent
render, and -- on a box with >= 2 GPUs -- N ranks over NCCL == one GPU, bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _acts(rng, batch):
    import torch
    return torch.from_numpy(rng.randint(0, 18, size=batch).astype(np.int32)).cuda()


@pytest.mark.parametrize('env_id', ['MoveToCorner-Demo-LoRes4E-v0', 'MatchRegions-Demo-LoRes4A-v0',
                                    'FindDupe-Demo-LoResStack-v0'])
def test_newest_frame_output_and_external_obs_slice(built, env_id):
    """The newest-frame send buffer equals the last 3 channels of the stack, and a handle
    bound to a slice of a larger tensor writes the same observation as a plain handle."""
    import torch
    import magical_b200 as magical
    B, pad = 24, 5
    ref = magical.make_vec(env_id, B, auto_reset=True)
    env = magical.make_vec(env_id, B, auto_reset=True)
    two = ref.obs.dim() == 5
    big_shape = ((2, B + 2 * pad, 96, 96, 12) if two else (B + 2 * pad, 96, 96, 12))
    big = torch.full(big_shape, 77, dtype=torch.uint8, device='cuda')
    view = big[:, pad:pad + B] if two else big[pad:pad + B]
    env.bind_obs(view)
    newest = torch.zeros(env.newest_shape(), dtype=torch.uint8, device='cuda')
    assert tuple(newest.shape) == ((2 if two else 1), B, 96, 96, 3)
    env.bind_newest(newest)
    ref.reset()
    env.reset()
    rng = np.random.RandomState(0)
    for t in range(ref.max_episode_steps + 6):     # through one auto-reset
        a = _acts(rng, B)
        o_ref = ref.step(a)[0]
        o = env.step(a)[0]
        assert torch.equal(o, o_ref), t
        want = o_ref[..., 9:12] if two else o_ref[None][..., 9:12]
        assert torch.equal(newest, want), t
    # nothing outside the slice was touched
    if two:
        assert int((big[:, :pad] != 77).sum()) == 0 and int((big[:, pad + B:] != 77).sum()) == 0
    else:
        assert int((big[:pad] != 77).sum()) == 0 and int((big[pad + B:] != 77).sum()) == 0
    ref.close()
    env.close()


def _ref_push(stacks, newest, fresh, first, count, shard, rank_stride):
    import torch
    fb = 96 * 96 * 3
    out = stacks.clone()
    for e in range(first, first + count):
        off = (e // shard) * rank_stride + (e % shard) * fb
        fr = newest[off:off + fb].view(96, 96, 3)
        out[e] = fr.repeat(1, 1, 4) if (fresh is not None and int(fresh[e])) else torch.cat([stacks[e][..., 3:], fr], -1)
    return out


def test_stack_push_kernel_matches_flatten_frame_stack(built):
    import torch
    from magical_b200 import dist as mdist
    g = torch.Generator(device='cuda').manual_seed(5)
    n, shard, views = 14, 7, 2          # 2 "ranks" of 7 envs, two views per rank in the gathered buffer
    fb = 96 * 96 * 3
    stacks = torch.randint(0, 256, (n, 96, 96, 12), dtype=torch.uint8, device='cuda', generator=g)
    recv = torch.randint(0, 256, (2 * views * shard * fb,), dtype=torch.uint8, device='cuda', generator=g)
    fresh = (torch.rand(n, device='cuda', generator=g) < 0.3).to(torch.uint8)
    for v in range(views):
        newest = recv[v * shard * fb:]
        for first, count, fr in ((0, n, fresh), (3, 6, fresh), (7, 7, None), (0, 0, fresh)):
            want = _ref_push(stacks, newest, fr, first, count, shard, views * shard * fb)
            got = stacks.clone()
            mdist.cuda_stack_push(got, newest, fr, first, count, shard, views * shard * fb)
            torch.cuda.synchronize()
            assert torch.equal(got, want), (v, first, count)


@pytest.mark.parametrize('env_id', ['MoveToCorner-Demo-LoRes4E-v0', 'MoveToCorner-Demo-LoRes3EA-v0',
                                    'MoveToCorner-Demo-LoResCHW4E-v0', 'MoveToCorner-Demo-LoResStack-v0'])
def test_render_is_idempotent_in_stacked_modes(built, env_id):
    """render() must not advance the frame stack (the reference's render() has no
    effect on the next observation), and step == step_physics + step_render."""
    import torch
    import magical_b200 as magical
    B = 6
    a_env = magical.make_vec(env_id, B, auto_reset=True)
    b_env = magical.make_vec(env_id, B, auto_reset=True)
    a_env.reset()
    b_env.reset()
    rng = np.random.RandomState(2)
    for t in range(12):
        a = _acts(rng, B)
        oa = a_env.step(a)[0].clone()
        b_env.step_physics(a)
        ob = b_env.step_render()
        assert torch.equal(oa, ob), t
        for _ in range(3):
            assert torch.equal(b_env.render(), oa), t
    a_env.close()
    b_env.close()


# ----------------------------------------------------------------- N ranks == 1 GPU
def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, env_id, total, steps, pipeline, transport, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import magical_b200 as magical
    from magical_b200 import dist as mdist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    # every rank builds the same scene pool (same seed) so that shards are slices of one global batch
    env = mdist.ShardedVecEnv(
        lambda n: magical.make_vec(env_id, n, device=rank, auto_reset=True, seed=11, alloc_obs=False,
                                   keep_scene=True),
        total, rank, world, gather_obs='newest', pipeline=pipeline, transport=transport)
    assert env.transport == transport
    n_sc = env.local.n_scenes
    if n_sc > 1:
        ids = np.arange(env.start, env.stop) % n_sc
        env.local.reset(scene_ids=ids)
        env._full_gather()
    else:
        env.reset()
    rng = np.random.RandomState(3)
    outs = []
    for t in range(steps):
        acts = torch.from_numpy(rng.randint(0, 18, size=total).astype(np.int32)).cuda()
        obs, rew, done, info = env.step(acts)
        env.wait_obs()
        torch.cuda.synchronize()
        if t % 7 == 0 or t >= steps - 3 or 38 <= t <= 41:   # 39: the MoveToRegion episodes end (auto-reset)
            outs.append((t, obs.cpu().numpy().copy(), rew.cpu().numpy().copy(), done.cpu().numpy().copy(),
                         info['eval_score'].cpu().numpy().copy()))
    np.savez(os.path.join(out_dir, f'rank{rank}.npz'), ts=np.array([o[0] for o in outs]),
             **{f'{k}{i}': o[j] for i, o in enumerate(outs) for j, k in ((1, 'obs'), (2, 'rew'), (3, 'done'), (4, 'score'))})
    assert env.local.overflow_count() == 0
    env.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('env_id,total,steps,pipeline,transport', [
    ('MoveToRegion-Demo-LoRes4E-v0', 64, 48, False, 'p2p'),   # 40-step episodes: one auto-reset inside
    ('MoveToRegion-Demo-LoRes4E-v0', 64, 48, True, 'p2p'),
    ('MatchRegions-TestAll-LoResStack-v0', 48, 30, True, 'p2p'),   # config 4's layout: two views, scene pool
    ('MoveToRegion-Demo-LoRes4E-v0', 64, 48, True, 'nccl'),
    ('MatchRegions-TestAll-LoResStack-v0', 48, 30, False, 'nccl'),
])
def test_n_gpu_equals_one_gpu_bit_for_bit(built, tmp_path, env_id, total, steps, pipeline, transport):
    """p2p: NVLink peer memory, frames read by the stack-rebuild kernel from the owners' buffers;
    nccl: ncclAllGather + k_stack_push.  Either way every rank must hold the single-GPU result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs >= 2 GPUs')
    import torch.multiprocessing as mp
    import magical_b200 as magical
    world = 2
    mp.spawn(_rank_main, args=(world, _free_port(), env_id, total, steps, pipeline, transport, str(tmp_path)),
             nprocs=world, join=True)
    # the same global batch on ONE GPU
    venv = magical.make_vec(env_id, total, device=0, auto_reset=True, seed=11, keep_scene=True)
    if venv.n_scenes > 1:
        venv.reset(scene_ids=np.arange(total) % venv.n_scenes)
    else:
        venv.reset()
    rng = np.random.RandomState(3)
    ranks = [np.load(os.path.join(str(tmp_path), f'rank{r}.npz')) for r in range(world)]
    ts = list(ranks[0]['ts'])
    for t in range(steps):
        acts = torch.from_numpy(rng.randint(0, 18, size=total).astype(np.int32)).cuda()
        obs, rew, done, info = venv.step(acts)
        if t in ts:
            i = ts.index(t)
            for r in ranks:
                assert np.array_equal(r[f'obs{i}'], obs.cpu().numpy()), (t, 'obs')
                assert np.array_equal(r[f'rew{i}'], rew.cpu().numpy()), (t, 'rew')
                assert np.array_equal(r[f'done{i}'], done.cpu().numpy()), (t, 'done')
                assert np.array_equal(r[f'score{i}'], info['eval_score'].cpu().numpy()), (t, 'score')
    assert any(bool(ranks[0][f'done{i}'].any()) for i in range(len(ts))) or steps < 40
    venv.close()
