"""
Host-side placement for records that start in host memory: bind the calling process to the CPUs of the NUMA node its
GPU hangs off, so that pinned staging buffers allocated afterwards are first-touched on that node and every rank of a
multi-GPU job copies over its own PCIe root complex instead of crossing the socket interconnect.

Linux sysfs only (``/sys/bus/pci/devices/<bus id>/numa_node``, ``/sys/devices/system/node/node<k>/cpulist``); where the
platform does not say (virtual machines report -1) nothing is changed and the reason is returned.
"""
import os


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(pci_bus_id):
    """NUMA node of a PCI device ('0000:1b:00.0'; the 8-digit domain nvidia-smi prints is accepted), or None."""
    bus = pci_bus_id.strip().lower()
    dom, _, rest = bus.partition(":")
    if len(dom) > 4:
        bus = dom[-4:] + ":" + rest
    try:
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
    except (OSError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPUs of the NUMA node of CUDA device ``device_index`` (call it before pinned
    buffers are allocated).  Returns a dict {node, cpus, bound, why} describing what was done."""
    info = {"node": None, "cpus": None, "bound": False, "why": ""}
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        pci = f"{dom:04x}:{bus:02x}:{dev:02x}.0"
    except Exception as exc:                     # noqa: BLE001 -- any failure means "leave the placement alone"
        info["why"] = f"no PCI id: {exc}"
        return info
    node = gpu_numa_node(pci)
    if node is None:
        info["why"] = f"{pci}: the platform reports no NUMA node"
        return info
    try:
        cpus = _parse_cpulist(open(f"/sys/devices/system/node/node{node}/cpulist").read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            info.update(node=node, why="no allowed CPU on that node")
            return info
        os.sched_setaffinity(0, use)
        info.update(node=node, cpus=len(use), bound=True)
    except (OSError, ValueError, AttributeError) as exc:
        info.update(node=node, why=str(exc))
    return info
