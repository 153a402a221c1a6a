"""Input-sharded MIMO (SURVEY.md 8e, C5 "NVLink mixdown"): rank g holds a shard of the INPUTS and every output of the
matrix, the partial output spectra are summed over the ranks (one reduce-scatter per call) and rank g converts its
shard of the OUTPUTS.

  * CPU (gloo, world_size 2): the host logic -- input shards from bbx_shard_range, partial results summed with a
    collective, output shards gathered -- with the CPU oracle standing in for the device.  Sum order differs from the
    single engine, so the comparison uses the path's tolerance (SNR >= 110 dB, max-abs <= 1e-5 x peak).
  * GPU, one device: the sharded code path of libbbx (gather kernel -> ncclReduceScatter over a 1-rank communicator
    -> strided inverse transforms) against the oracle.
  * GPU, two devices (skipped on a one-GPU box): two processes, NCCL over NVLink, against the oracle.
"""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

B, L, NIN, NOUT, NBLK, T = 128, 700, 6, 4, 32, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _paths():
    for p in (ROOT, HERE):
        if p not in sys.path:
            sys.path.insert(0, p)


def _full_oracle(nin=NIN, nout=NOUT):
    _paths()
    import cpulibs as cl
    from convkit import OracleDriver, interleave, make_ir, make_noise, run_float
    P = -(-L // B)
    o = OracleDriver(B, P, nin, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=T)
    for oo in range(nout):
        for i in range(nin):
            o.select(oo * nin + i, o.filter(make_ir(7000 + 64 * oo + i, L)))
    x = interleave([make_noise(7100 + i, NBLK * B) for i in range(nin)])
    return x, run_float(o, x, T * B)


# ------------------------------------------------------------------------------------------------
# CPU: host logic with gloo
# ------------------------------------------------------------------------------------------------
def _cpu_worker(rank, world, port, ret):
    _paths()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bbcat_dsp_b200 as bbx
    import cpulibs as cl
    from convkit import OracleDriver, interleave, make_ir, make_noise, run_float
    P = -(-L // B)
    i0, ni = bbx.shard_range(NIN, rank, world)     # this rank's inputs
    o0, no = bbx.shard_range(NOUT, rank, world)    # this rank's outputs after the reduce-scatter
    o = OracleDriver(B, P, ni, n_outputs=NOUT, mode=cl.MODE_MIMO, max_blocks=T)
    for oo in range(NOUT):
        for i in range(ni):
            o.select(oo * ni + i, o.filter(make_ir(7000 + 64 * oo + i0 + i, L)))
    x = interleave([make_noise(7100 + i0 + i, NBLK * B) for i in range(ni)])
    part = run_float(o, x, T * B)                  # partial outputs [frames][NOUT] (linear: time or frequency domain)
    # reduce-scatter over the output axis (gloo has no reduce_scatter: all_reduce + slice is the same arithmetic)
    t = torch.from_numpy(np.ascontiguousarray(part.T)).clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    mine = t[o0:o0 + no].numpy()
    bufs = [torch.zeros((NOUT // world, NBLK * B), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(np.ascontiguousarray(mine)))
    if rank == 0:
        ret["y"] = np.concatenate([b.numpy() for b in bufs], axis=0).T
    dist.destroy_process_group()


def test_input_sharded_mimo_host_logic_gloo_world2():
    _paths()
    from parity import assert_float_parity
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cpu_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    _, y_full = _full_oracle()
    y = ret["y"]
    assert y.shape == y_full.shape
    for oo in range(NOUT):
        assert_float_parity(y[:, oo], y_full[:, oo], "sharded sum out %d" % oo)


# ------------------------------------------------------------------------------------------------
# output-sharded layout (the default MIMO split): rank g holds EVERY input and the rows
# [g n_out / world, (g+1) n_out / world) of the matrix; no collective, the ranks' PCM outputs are just concatenated
# ------------------------------------------------------------------------------------------------
def _output_shard_engine(make_driver, rank, world, nin=NIN, nout=NOUT, **kw):
    _paths()
    import bbcat_dsp_b200 as bbx
    import cpulibs as cl
    from convkit import make_ir
    P = -(-L // B)
    o0, no = bbx.shard_range(nout, rank, world)
    d = make_driver(B, P, nin, n_outputs=no, mode=cl.MODE_MIMO, max_blocks=T, **kw)
    for oo in range(no):
        for i in range(nin):
            d.select(oo * nin + i, d.filter(make_ir(7000 + 64 * (o0 + oo) + i, L)))
    return d


def _cpu_worker_outputs(rank, world, port, ret):
    _paths()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from convkit import OracleDriver, interleave, make_noise, run_float
    o = _output_shard_engine(OracleDriver, rank, world)
    x = interleave([make_noise(7100 + i, NBLK * B) for i in range(NIN)])  # every rank sees all inputs
    mine = run_float(o, x, T * B)  # [frames][NOUT / world]
    bufs = [torch.zeros((NBLK * B, NOUT // world), dtype=torch.float32) for _ in range(world)]
    dist.all_gather(bufs, torch.from_numpy(np.ascontiguousarray(mine)))
    if rank == 0:
        ret["y"] = np.concatenate([b.numpy() for b in bufs], axis=1)
    dist.destroy_process_group()


def test_output_sharded_mimo_host_logic_gloo_world2():
    """two ranks, each with all inputs and half of the outputs; no reduction.  Every output is summed exactly as in the
    unsharded convolver (same inputs, same order), so the oracle's shards equal the full oracle bit for bit."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_cpu_worker_outputs, args=(2, _free_port(), ret), nprocs=2, join=True)
    _, y_full = _full_oracle()
    y = ret["y"]
    assert y.shape == y_full.shape
    assert np.array_equal(y.view(np.uint32), y_full.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("tensor_off", [0, 1])
def test_output_sharded_mimo_product_vs_oracle(bbx, tensor_off):
    """the product's output shards (two engines, each all inputs x half of the outputs; tensor-core and SIMT MAC) against the
    full oracle.  Tolerance class: the MAC plan of a shard cuts the row space differently from the full engine."""
    from convkit import GpuDriver, interleave, make_noise, run_float
    from parity import assert_float_parity
    x = interleave([make_noise(7100 + i, NBLK * B) for i in range(NIN)])
    parts = []
    for rank in range(2):
        g = _output_shard_engine(lambda *a, **kw: GpuDriver(bbx, *a, **kw), rank, 2, mimo_tensor=tensor_off)
        parts.append(run_float(g, x, [T * B, 3 * B, (T - 3) * B]))
        n_tc, status = g.eng.tensor_status()
        assert status == 0 and (n_tc > 0) == (not tensor_off)
        g.close()
    y = np.concatenate(parts, axis=1)
    _, y_full = _full_oracle()
    for oo in range(NOUT):
        assert_float_parity(y[:, oo], y_full[:, oo], "output-sharded out %d" % oo)


# ------------------------------------------------------------------------------------------------
# GPU: the product path
# ------------------------------------------------------------------------------------------------
def _gpu_run(bbx, rank, world, comm, tensor_off, nin=NIN, nout=NOUT, peer=False):
    """one rank of the sharded engine; returns [frames][nout / world].  peer: the partial spectra travel as peer-memory
    stores (CUDA IPC buffers, handles gathered over the gloo group) instead of an NCCL reduce-scatter."""
    import cpulibs as cl
    from convkit import GpuDriver, interleave, make_ir, make_noise, run_float
    P = -(-L // B)
    i0, ni = bbx.shard_range(nin, rank, world)
    g = GpuDriver(bbx, B, P, ni, n_outputs=nout, mode=cl.MODE_MIMO, max_blocks=T, mimo_tensor=tensor_off,
                  mimo_shard_world=world, mimo_shard_rank=rank, device=rank if world > 1 else 0)
    if peer:
        handles = [None] * world
        dist.all_gather_object(handles, g.eng.PeerExport())
        g.eng.PeerAttach(handles)
    else:
        g.eng.SetComm(comm)
    for oo in range(nout):
        for i in range(ni):
            g.select(oo * ni + i, g.filter(make_ir(7000 + 64 * oo + i0 + i, L)))
    x = interleave([make_noise(7100 + i0 + i, NBLK * B) for i in range(ni)])
    y = run_float(g, x, [T * B, 3 * B, (T - 3) * B])  # tensor-core call, SIMT call, SIMT/TC call
    st = g.eng.tensor_status()
    if peer:
        dist.barrier()  # nobody frees a receive buffer that the others still have mapped
    g.close()
    return y, st


@pytest.mark.gpu
@pytest.mark.parametrize("tensor_off", [0, 1])
def test_sharded_path_one_rank_vs_oracle(bbx, tensor_off):
    from parity import assert_float_parity
    if not bbx.lib().bbx_comm_available():
        pytest.skip("libnccl.so.2 not loadable")
    comm = bbx.Comm(1, 0, bbx.comm_unique_id(), device=0)
    y, (n_tc, status) = _gpu_run(bbx, 0, 1, comm, tensor_off)
    comm.close()
    assert status == 0 and (n_tc > 0) == (not tensor_off)
    _, y_full = _full_oracle()
    for oo in range(NOUT):
        assert_float_parity(y[:, oo], y_full[:, oo], "1-rank sharded path out %d" % oo)


def _gpu_worker(rank, world, port, ret, nin=NIN, nout=NOUT, peer=False):
    _paths()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bbcat_dsp_b200 as bbx
    import torch
    torch.cuda.set_device(rank)
    if peer:
        y, st = _gpu_run(bbx, rank, world, None, 0, nin, nout, peer=True)
        dist.barrier()  # every rank has finished reading the others' stores before anybody runs again
        y2, _ = _gpu_run(bbx, rank, world, None, 0, nin, nout, peer=True)
        assert np.array_equal(y.view(np.uint32), y2.view(np.uint32)), "peer mixdown is not bit-reproducible"
    else:
        uid = [bbx.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        comm = bbx.Comm(world, rank, uid[0], device=rank)
        y, st = _gpu_run(bbx, rank, world, comm, 0, nin, nout)
        comm.close()
    ret[rank] = (y, st)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_input_sharded_mimo_two_gpus_vs_oracle(bbx):
    from parity import assert_float_parity
    if bbx.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gpu_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
    _, y_full = _full_oracle()
    y = np.concatenate([ret[0][0], ret[1][0]], axis=1)
    assert ret[0][1][1] == 0 and ret[1][1][1] == 0 and ret[0][1][0] > 0
    for oo in range(NOUT):
        assert_float_parity(y[:, oo], y_full[:, oo], "2-GPU sharded MIMO out %d" % oo)


@pytest.mark.gpu
def test_input_sharded_mimo_all_gpus_vs_oracle(bbx):
    """The same on every GPU of the box (4 or 8 ranks: 16 inputs and 8 outputs of the matrix sharded over them)."""
    from parity import assert_float_parity
    world = 8 if bbx.device_count() >= 8 else 4
    if bbx.device_count() < world:
        pytest.skip("needs four or eight GPUs (gpurun --gpus 4 / 8)")
    nin, nout = 16, 8
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), ret, nin, nout), nprocs=world, join=True)
    _, y_full = _full_oracle(nin, nout)
    y = np.concatenate([ret[r][0] for r in range(world)], axis=1)
    assert all(ret[r][1][1] == 0 for r in range(world)) and ret[0][1][0] > 0
    for oo in range(nout):
        assert_float_parity(y[:, oo], y_full[:, oo], "%d-GPU sharded MIMO out %d" % (world, oo))


@pytest.mark.gpu
def test_peer_mixdown_two_gpus_vs_oracle(bbx):
    """Input-sharded MIMO with the peer-memory mixdown (NVLink stores into CUDA IPC buffers + epoch flags, no NCCL):
    two processes on two GPUs against the oracle, and bit-reproducible from run to run (fixed rank-order sums)."""
    from parity import assert_float_parity
    if bbx.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gpu_worker, args=(2, _free_port(), ret, NIN, NOUT, True), nprocs=2, join=True)
    _, y_full = _full_oracle()
    y = np.concatenate([ret[0][0], ret[1][0]], axis=1)
    assert ret[0][1][1] == 0 and ret[1][1][1] == 0 and ret[0][1][0] > 0
    for oo in range(NOUT):
        assert_float_parity(y[:, oo], y_full[:, oo], "2-GPU peer mixdown out %d" % oo)


@pytest.mark.gpu
def test_peer_mixdown_all_gpus_vs_oracle(bbx):
    from parity import assert_float_parity
    world = 8 if bbx.device_count() >= 8 else 4
    if bbx.device_count() < world:
        pytest.skip("needs four or eight GPUs (gpurun --gpus 4 / 8)")
    nin, nout = 16, 8
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gpu_worker, args=(world, _free_port(), ret, nin, nout, True), nprocs=world, join=True)
    _, y_full = _full_oracle(nin, nout)
    y = np.concatenate([ret[r][0] for r in range(world)], axis=1)
    assert all(ret[r][1][1] == 0 for r in range(world)) and ret[0][1][0] > 0
    for oo in range(nout):
        assert_float_parity(y[:, oo], y_full[:, oo], "%d-GPU peer mixdown out %d" % (world, oo))
