"""CPU: the N > 1 host logic (usdm_b200/parallel.py) over world_size-2 `gloo`.

The data path has no collective (utterances are independent); what is tested here is the sharding,
micro-batching and the single all-gather of the output mels.  The per-shard `generate_fn` is a deterministic stand-in
(the CUDA engine needs a GPU); the result must be identical to running the stand-in over the whole batch on one rank.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from util import ROOT  # noqa: F401  (sys.path)


def _fake_generate(x, cond, lengths, **kw):
    """Stand-in with the Voicebox.generate contract: [n,T] ids, [n,M,T] cond, [n] lengths -> [n,M,T], zeros past length."""
    n, M, T = cond.shape
    out = cond * 2.0 + x.to(torch.float32).unsqueeze(1) * 1e-3 + float(kw.get("n_timesteps", 0))
    mask = (torch.arange(T).unsqueeze(0) < lengths.unsqueeze(1)).unsqueeze(1)
    return out * mask


def _fake_vocode(mel):
    """Stand-in with the BigVGAN.forward contract ([n,M,t] -> [n,1,t*4]) whose output depends on the utterance's own length
    (like a conv stack's edge behaviour), so batching at a padded length would be detected."""
    n, M, t = mel.shape
    return (mel.sum(1, keepdim=True).repeat_interleave(4, dim=2) + float(t)).tanh()


def _inputs():
    g = torch.Generator().manual_seed(7)
    lens = torch.tensor([50, 7, 33, 12, 64, 1, 20, 41, 5])
    N, M, T = len(lens), 4, int(lens.max())
    x = torch.randint(0, 100, (N, T), generator=g)
    cond = torch.randn(N, M, T, generator=g)
    return x, cond, lens


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from usdm_b200 import parallel
        x, cond, lens = _inputs()
        full = parallel.generate_sharded(_fake_generate, x, cond, lens, max_tokens=150, n_timesteps=3)
        wav = parallel.vocode_sharded(_fake_vocode, full, lens, hop=4, max_frames=60)
        q.put((rank, (full, wav)))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(300)
def test_generate_sharded_world2_gloo_equals_single_rank():
    from usdm_b200 import parallel
    x, cond, lens = _inputs()
    want = _fake_generate(x, cond, lens, n_timesteps=3)
    single = parallel.generate_sharded(_fake_generate, x, cond, lens, max_tokens=150, n_timesteps=3)   # world 1, no process group
    assert torch.equal(single, want)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=240) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want_wav = torch.zeros(len(lens), 1, int(lens.max()) * 4)
    for u, t in enumerate(lens.tolist()):
        want_wav[u, :, :t * 4] = _fake_vocode(want[u:u + 1, :, :t])[0]
    assert torch.equal(parallel.vocode_sharded(_fake_vocode, want, lens, hop=4, max_frames=60), want_wav)      # world 1
    for r in range(2):
        assert torch.equal(got[r][0], want), f"rank {r} mel differs"
        assert torch.equal(got[r][1], want_wav), f"rank {r} waveform differs"


def test_shards_are_balanced_deterministic_and_cover_everything():
    from usdm_b200 import parallel
    rng = torch.Generator().manual_seed(1234)
    lens = torch.randint(172, 1723, (512,), generator=rng).tolist()      # configs[4]-like: 2-20 s
    for world in (1, 2, 4, 8):
        shards = parallel.shard_utterances(lens, world)
        assert sorted(i for s in shards for i in s) == list(range(512))
        assert shards == parallel.shard_utterances(lens, world)
        loads = [sum(parallel.utterance_cost(lens[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.01                    # LPT keeps the ranks within 1 %
    mbs = parallel.micro_batches(shards[0], lens, max_tokens=40000)
    assert sorted(i for b in mbs for i in b) == sorted(shards[0])
    for b in mbs:
        assert len(b) == 1 or sum(2 * (lens[i] + 1) for i in b) <= 40000
