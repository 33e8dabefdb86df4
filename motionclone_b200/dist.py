"""Multi-GPU plumbing: independent samples shard across ranks; ONE broadcast of the packed motion representation.

The reference is single-process / single-GPU. Samples (JSONL lines, t2v_video_sample.py:75-105) are
independent, so the path shards with no data-path collective; the only shared state is the reference clip's motion
representation (6 x (fp16 values + uint8 indices), 590 KB at 16x512x512), which rank 0 extracts once and broadcasts
as a single contiguous byte buffer (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the process group when world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend=backend or ("nccl" if torch.cuda.is_available() else "gloo"),
                                rank=rank, world_size=world)
    return rank, world, local


def shard_samples(n_samples: int, rank: int, world: int) -> List[int]:
    """Round-robin: rank r takes samples r, r+W, ...."""
    return list(range(rank, n_samples, world))


def pack_representation(rep: Dict[str, Sequence[torch.Tensor]]) -> Tuple[torch.Tensor, list]:
    """-> (uint8 buffer, manifest [(name, val_shape, idx_shape)]); values fp16, indices uint8, module order kept."""
    chunks, manifest = [], []
    for name, (val, idx) in rep.items():
        v = val.detach().to(torch.float16).contiguous()
        i = idx.detach().to(torch.uint8).contiguous()
        manifest.append((name, tuple(v.shape), tuple(i.shape)))
        chunks += [v.view(torch.uint8).reshape(-1), i.reshape(-1)]
    return torch.cat(chunks), manifest


def unpack_representation(buf: torch.Tensor, manifest: list) -> Dict[str, List[torch.Tensor]]:
    out, off = {}, 0
    for name, vshape, ishape in manifest:
        nv = 2 * int(torch.Size(vshape).numel())
        ni = int(torch.Size(ishape).numel())
        val = buf[off:off + nv].clone().view(torch.float16).reshape(vshape)
        off += nv
        idx = buf[off:off + ni].clone().reshape(ishape)
        off += ni
        out[name] = [val, idx]
    return out


def broadcast_representation(rep, device, src: int = 0):
    """B1: rank `src` passes its representation, the others pass None; everyone gets the dict."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return rep
    rank = dist.get_rank()
    meta = [None]
    buf = None
    if rank == src:
        buf, manifest = pack_representation(rep)
        meta = [(manifest, buf.numel())]
    dist.broadcast_object_list(meta, src=src)  # shapes only (host side, a few hundred bytes)
    manifest, nbytes = meta[0]
    if rank != src:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    else:
        buf = buf.to(device)
    dist.broadcast(buf, src=src)  # the one data collective
    return unpack_representation(buf, manifest)
