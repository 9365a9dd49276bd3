"""Bind the calling process to the NUMA node its GPU hangs off (host-side plumbing for the pinned-buffer pipeline).

Pinned pages are placed on the node of the thread that allocates them; on a two-socket box GPUs 4-7 sit on node 1, so
a rank that allocates its staging buffers from node 0 pays a cross-socket hop on every H2D / D2H copy (SURVEY 8f
rank 2: "at GPU speeds host decode + PCIe dominates end-to-end").  Everything here is best effort: on a box without
sysfs NUMA information nothing changes and ``None`` is returned.
"""
import ctypes
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def parse_cpulist(text):
    """'0-3,8,10-11' -> [0, 1, 2, 3, 8, 10, 11]"""
    cpus = []
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_numa_node(index):
    """NUMA node of CUDA device ``index`` (by PCI bus id), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [v for v in vis.split(",") if v.strip() != ""]
            if index < len(ids) and ids[index].strip().isdigit():
                phys = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
    except Exception:
        return None
    # NVML prints an 8-digit domain ("00000000:1B:00.0"), sysfs uses 4 ("0000:1b:00.0")
    dom, rest = bus.split(":", 1)
    node = _read("/sys/bus/pci/devices/%s:%s/numa_node" % (dom[-4:].lower(), rest.lower()))
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def node_cpus(node):
    return parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % node))


def bind_to_gpu(index):
    """Restrict this process to the CPUs of the GPU's NUMA node and prefer that node for new pages.
    Returns {"node": n, "cpus": k} or None when the topology is unknown."""
    node = gpu_numa_node(index)
    if node is None:
        return None
    cpus = set(node_cpus(node))
    try:
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
    except (AttributeError, OSError):
        return None
    try:        # set_mempolicy(MPOL_PREFERRED, {node}): first-touch already does this for a bound thread; belt and braces
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] |= 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        libc.syscall(238, 1, mask, 16 * 8 * ctypes.sizeof(ctypes.c_ulong))     # x86_64 __NR_set_mempolicy
    except Exception:
        pass
    return {"node": node, "cpus": len(cpus)}
