"""What caps the end-to-end (host-buffer) path at 8 GPUs?  Every rank copies from pinned host memory to its GPU at the same
time; the pinned buffer is allocated (a) as cudaHostAlloc places it, (b) after set_mempolicy(MPOL_BIND) to the NUMA node of
the rank's GPU.  Prints the box's topology first.   torchrun --nproc-per-node 8 scripts/h2d_numa_check.py  (or python ... for one GPU)"""
import ctypes as C
import glob
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                                   # noqa: E402
import torch.distributed as dist              # noqa: E402
from manifoldem_python_b200 import _lib       # noqa: E402

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
if world > 1:
    torch.cuda.set_device(local)
    dist.init_process_group('nccl')


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return repr(e)


def gpu_numa_node(index):
    bdf = sh('nvidia-smi -i %d --query-gpu=pci.bus_id --format=csv,noheader' % index).lower()
    if bdf.startswith('00000000:'):
        bdf = '0000:' + bdf[9:]
    try:
        return int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read()), bdf
    except Exception:
        return -1, bdf


if rank == 0:
    print('cpus visible', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
    print('nodes', [os.path.basename(p) for p in glob.glob('/sys/devices/system/node/node[0-9]*')])
    print(sh("grep -E 'Cpus_allowed_list|Mems_allowed_list' /proc/self/status"))
    print(sh('nvidia-smi topo -m | head -14'))
    for n in glob.glob('/sys/devices/system/node/node[0-9]*/meminfo'):
        print(sh("grep -E 'MemTotal|MemFree' %s" % n))
libc = C.CDLL(None, use_errno=True)
node, bdf = gpu_numa_node(local)
lib = _lib.load()
ctx = _lib.Context(local)
n = 512 << 20
d = _lib.DeviceArray(ctx, (n,), np.uint8)


def run(tag, bind):
    ok = None
    if bind and node >= 0:
        mask = C.c_ulong(1 << node)
        ok = libc.syscall(238, 2, C.byref(mask), 65)          # set_mempolicy(MPOL_BIND, {node}) on x86-64
    h = _lib.PinnedArray((n,), np.uint8)
    h.array[::4096] = 1
    if bind and node >= 0:
        libc.syscall(238, 0, None, 0)                          # MPOL_DEFAULT
    _lib.check(lib.mem_copy_h2d(ctx.handle, d.ptr, h.ptr, n))
    barrier()
    t0 = time.perf_counter()
    for _ in range(8):
        _lib.check(lib.mem_copy_h2d(ctx.handle, d.ptr, h.ptr, n))
    dt = time.perf_counter() - t0
    barrier()
    h.free()
    gbs = 8 * n / dt / 1e9
    t = torch.tensor([gbs], device='cuda:%d' % local, dtype=torch.float64)
    allv = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allv, t)
    else:
        allv = [t]
    if rank == 0:
        v = [float(x) for x in allv]
        print('%-28s per-GPU GB/s: %s   sum %.0f' % (tag, ' '.join('%.1f' % x for x in v), sum(v)))
    return ok


for rep in range(2):
    run('default placement', False)
    ok = run('bound to the GPU\'s NUMA node', True)
print('rank %d gpu %s numa node %d set_mempolicy rc %s' % (rank, bdf, node, ok), flush=True)
if world > 1:
    dist.destroy_process_group()
