"""World-size-2 (and 3, ragged) runs of the sharding / all-gather plumbing on
CPU with the gloo backend; the per-rank 'renderer' is the CPU oracle so the
gathered result can be checked against a single-process render."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nerf_from_image_b200 import parallel
from fixtures import synthetic
from tests import helpers as Hh


def test_shard_range_covers_batch():
    for batch in (1, 2, 5, 16, 32, 33):
        for world in (1, 2, 3, 4, 8):
            spans = [parallel.shard_range(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and b >= a
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, out_dir, inplace=False):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    H, W, S = 8, 8, 8
    scene, cams = Hh.make_case('p3d_bbox', seed=5, batch=batch, plane_res=16)
    nt, nu = synthetic.make_noise(5, batch, H, W, S)

    def render_fn(planes, palette, c2w, focal, bbox, noise_t, noise_u, out=None):
        sc = dict(scene, planes=planes, palette=palette)
        cm = dict(c2w=c2w, focal=focal, center=None, bbox=bbox)
        o = Hh.run_oracle(sc, cm, H, W, S, noise_t, noise_u.reshape(-1, S))
        if out is not None:  # what fused_render(out=...) does: results land in the slices
            for dst, k in zip(out, ('rgb', 'depth', 'mask')):
                dst.copy_(o[k])
            return out
        return o['rgb'], o['depth'], o['mask']

    inputs = dict(planes=scene['planes'], palette=scene['palette'], c2w=cams['c2w'],
                  focal=cams['focal'], bbox=cams['bbox'], noise_t=nt,
                  noise_u=nu.view(batch, H * W, S))
    if inplace:
        rgb, depth, mask = parallel.render_sharded(render_fn, inputs, batch, inplace=True,
                                                   height=H, width=W)
    else:
        rgb, depth, mask = parallel.render_sharded(render_fn, inputs, batch)
    torch.save((rgb, depth, mask), os.path.join(out_dir, 'r%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,batch,inplace', [(2, 4, False), (3, 5, False), (2, 4, True),
                                                 (3, 6, True), (3, 5, True)])
def test_sharded_render_equals_single_process(tmp_path, world, batch, inplace):
    """packed exchange (incl. ragged 5 over 3), in-place exchange (equal shards), and the
    in-place request falling back to the packed form on a ragged batch"""
    port = _free_port()
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path), inplace), nprocs=world, join=True)
    H, W, S = 8, 8, 8
    scene, cams = Hh.make_case('p3d_bbox', seed=5, batch=batch, plane_res=16)
    nt, nu = synthetic.make_noise(5, batch, H, W, S)
    ref = Hh.run_oracle(scene, cams, H, W, S, nt, nu)
    for r in range(world):
        rgb, depth, mask = torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r))
        assert rgb.shape == ref['rgb'].shape
        # per-image arithmetic only: the gathered batch is bit-identical
        assert torch.equal(rgb, ref['rgb'])
        assert torch.equal(depth, ref['depth'])
        assert torch.equal(mask, ref['mask'])


def _row_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(1)
    B, H, W, S = 1, 40, 8, 8            # ONE image, more ranks than images: rows are split
    scene, cams = Hh.make_case('p3d_bbox', seed=5, batch=B, plane_res=16)
    nt, nu = synthetic.make_noise(5, B, H, W, S)

    def render_fn(r0, r1):
        nt_, nu_ = parallel.slice_rows(nt, nu, B, H, W, r0, r1)
        o = Hh.run_oracle(scene, cams, r1 - r0, W, S, nt_, nu_, rows=(r0, H),
                          global_near_far_fallback=False)
        return o['rgb'], o['depth'], o['mask']

    rgb, depth, mask = parallel.render_row_sharded(render_fn, H)
    torch.save((rgb, depth, mask), os.path.join(out_dir, 'r%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3, 7])
def test_row_sharded_render_equals_single_process(tmp_path, world):
    """Fewer images than ranks: every image's rows are split over the ranks
    in multiples of the 8-row tile height (40 rows: 24 + 16, 16 + 16 + 8, and five ranks of seven
    with rows), gathered back along the rows."""
    assert [parallel.row_range(40, 3, r) for r in range(3)] == [(0, 16), (16, 32), (32, 40)]
    assert [parallel.row_range(40, 7, r) for r in range(7)] == [(0, 8), (8, 16), (16, 24), (24, 32),
                                                                (32, 40), (40, 40), (40, 40)]
    mp.spawn(_row_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    B, H, W, S = 1, 40, 8, 8
    scene, cams = Hh.make_case('p3d_bbox', seed=5, batch=B, plane_res=16)
    nt, nu = synthetic.make_noise(5, B, H, W, S)
    ref = Hh.run_oracle(scene, cams, H, W, S, nt, nu, global_near_far_fallback=False)
    for r in range(world):
        rgb, depth, mask = torch.load(os.path.join(str(tmp_path), 'r%d.pt' % r))
        assert torch.equal(rgb, ref['rgb']) and torch.equal(depth, ref['depth'])
        assert torch.equal(mask, ref['mask'])


def _grad_worker(rank, world, port, out_dir):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)),
              torch.nn.Parameter(torch.zeros(1))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5.) * (rank + 1)
    # params[2] has no gradient on rank 0 (a leaf the rank's shard did not touch)
    if rank > 0:
        params[2].grad = torch.tensor([10.0 * rank])
    parallel.all_reduce_grads(params)
    torch.save([p.grad for p in params], os.path.join(out_dir, 'g%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_all_reduce(tmp_path):
    """GAN generator step (BASELINE config 4): decoder / beta / alpha gradients of the shards
    are summed with one bucketed all-reduce (run.py:636-644 DataParallel reduce)."""
    world = 3
    mp.spawn(_grad_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        g0, g1, g2 = torch.load(os.path.join(str(tmp_path), 'g%d.pt' % r))
        assert torch.equal(g0, torch.full((3, 4), 6.0))
        assert torch.equal(g1, torch.arange(5.) * 6)
        assert torch.equal(g2, torch.tensor([30.0]))
