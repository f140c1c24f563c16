"""Row-ownership sharding across GPUs: nodes (hence DOF rows 6p..6p+5) are
block-partitioned into contiguous ranges; each rank evaluates every element that touches an
owned node ("halo elements duplicated") and assembles its own CSR row block, so no collective
sits on the data path.  Host-side numpy only; the optional gathers use torch.distributed."""
import numpy as np


def node_ranges(nnodes, world_size):
    """Contiguous, balanced node ranges [(begin, end)] for every rank."""
    cuts = np.linspace(0, nnodes, world_size + 1).astype(np.int64)
    return [(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:])]


def elements_touching(conn, begin, end):
    """Indices of the elements with at least one node in [begin, end): the owned + halo set."""
    conn = np.asarray(conn)
    return np.nonzero(((conn >= begin) & (conn < end)).any(axis=1))[0]


def owned_elements(conn, begin, end):
    """Disjoint cover of the elements: an element belongs to the rank that owns its first node."""
    conn = np.asarray(conn)
    return np.nonzero((conn[:, 0] >= begin) & (conn[:, 0] < end))[0]


def shard_case(case, rank, world_size):
    """Sub-case for one rank: global node arrays, the elements touching its node range."""
    nnodes = case["ndof"] // 6
    begin, end = node_ranges(nnodes, world_size)[rank]
    sel = elements_touching(case["conn"], begin, end)
    sub = dict(case)
    sub["conn"] = np.asarray(case["conn"])[sel]
    for k in ("prop_id", "xmat", "vxy", "axes", "k", "hg", "K6ROT", "alpha"):
        v = case.get(k)
        if v is not None and np.ndim(v) >= 1 and np.shape(v)[0] == np.shape(case["conn"])[0]:
            sub[k] = np.asarray(v)[sel]
    sub["owned_nodes"] = (begin, end)
    sub["element_ids"] = sel
    return sub


def gather_row_blocks(local_indptr, local_nnz):
    """Optional NCCL/gloo step (never on the hot path): all-gather every rank's nnz so that a rank can
    offset its local indptr into the global CSR numbering.  Returns (global_row_offset_nnz, all_nnz)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return 0, [int(local_nnz)]
    dev = local_indptr.device if hasattr(local_indptr, "device") else "cpu"
    t = torch.tensor([int(local_nnz)], dtype=torch.int64, device=dev)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    sizes = [int(o.item()) for o in out]
    return sum(sizes[:dist.get_rank()]), sizes
