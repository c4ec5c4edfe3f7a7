"""Sharded multi-checkpoint zero-shot evaluation driver: the loop of the reference's evaluation scripts
(scripts/evaluate_domainnet_lso_openai.py:18-36 `get_data`, :214-228 the checkpoint loop; slurm/evaluate-clip.sh:131-133 runs it
for 33 checkpoints x 2 datasets) on top of the B200 path.

What changes against the reference loop, and why:
  * ONE model instance: every checkpoint is loaded INTO it (`load_checkpoint`), so device buffers, the towers' workspaces and
    their captured CUDA graphs survive from checkpoint to checkpoint (the engines refresh their operands in place);
  * the class-prompt classifier of a checkpoint is built in a few large truncated encode_text calls (OpenAIZeroShotClassifier);
  * images are sharded over the ranks batch by batch (rank r takes batches r, r + world, ...): weights and class prompts are
    replicated, no collective touches the data path, and the ONLY collective per (checkpoint, dataset) is the final
    all_reduce(SUM) of [top-1 hits, top-5 hits, samples];
  * uint8 batches are accepted as they come out of decode / resize / crop (ToTensor + Normalize run on the GPU).
Everything outside that loop (datasets, result serialisation, plotting) stays with the caller, as in the reference.
"""
from __future__ import annotations

from typing import Callable, Iterable, Optional, Sequence

import torch

__all__ = ["shard_batches", "get_data", "evaluate_model", "evaluate_checkpoints"]


def shard_batches(batches: Iterable, rank: int = 0, world_size: int = 1):
    """Batch i goes to rank i % world_size (strided: every rank sees the same mix of early and late samples)."""
    if not 0 <= rank < world_size:
        raise ValueError(f"rank {rank} outside [0, {world_size})")
    for i, batch in enumerate(batches):
        if i % world_size == rank:
            yield batch


@torch.inference_mode()
def get_data(clip, dataset, keys: Sequence[str], num_workers: int = 0, batch_size: int = 250, rank: int = 0, world_size: int = 1,
             input_dtype: Optional[torch.dtype] = None) -> dict:
    """Same contract as the reference's `get_data` (L2-normalised image features + the requested label columns), for THIS
    rank's shard of the dataset.  Batches are `(image, *labels)`; images may be uint8 pixels or float tensors (cast to the
    tower dtype like the reference's `.half()`)."""
    from torch.utils.data import DataLoader
    loader = DataLoader(dataset, batch_size=batch_size, num_workers=num_workers)
    clip.eval()
    data = {"img_feat": [], **{k: [] for k in keys}}
    for batch in shard_batches(loader, rank, world_size):
        img = batch[0].to("cuda", non_blocking=True)
        if img.dtype != torch.uint8 and input_dtype is not None:
            img = img.to(input_dtype)
        data["img_feat"].append(clip.encode_image(img, normalize=True))
        for i, key in enumerate(keys):
            data[key].append(batch[i + 1])
    return {k: (torch.cat(v) if v else torch.empty(0)) for k, v in data.items()}


def _reduce_hits(hits: torch.Tensor, world_size: int) -> torch.Tensor:
    if world_size > 1:
        import torch.distributed as dist
        dist.all_reduce(hits, op=dist.ReduceOp.SUM)          # the only collective of the evaluation
    return hits


@torch.inference_mode()
def evaluate_model(clip, tokenizer, dataset, classnames, domain_invariant: bool = False, batch_size: int = 250, num_workers: int = 0,
                   rank: int = 0, world_size: int = 1, topk: int = 5, input_dtype: Optional[torch.dtype] = None,
                   classifier_factory: Optional[Callable] = None) -> dict:
    """Top-1 / top-k accuracy of one checkpoint on one dataset, summed over all ranks.  `dataset` yields `(image, label)`."""
    if classifier_factory is None:
        from .zero_shot import OpenAIZeroShotClassifier
        classifier_factory = OpenAIZeroShotClassifier
    zs = classifier_factory(clip, tokenizer, classnames, domain_invariant)
    shard = get_data(clip, dataset, ["clss"], num_workers, batch_size, rank, world_size, input_dtype)
    dev = zs.prompt_feat.device
    hits = torch.zeros(3, dtype=torch.int64, device=dev)
    if shard["img_feat"].numel():
        labels = shard["clss"].to(dev).long()
        k = min(topk, len(classnames), 8)
        idx = zs.predict_topk_from_features(shard["img_feat"], k)["pred"]
        hits[0] = (idx[:, 0] == labels).sum()
        hits[1] = (idx == labels[:, None]).any(dim=1).sum()
        hits[2] = labels.numel()
    hits = _reduce_hits(hits, world_size).cpu()
    n = max(int(hits[2]), 1)
    return {"top1": int(hits[0]) / n, f"top{topk}": int(hits[1]) / n, "num-samples": int(hits[2])}


def evaluate_checkpoints(model_name: str, ckpt_files: Sequence[str], datasets: dict, classnames: dict, tokenizer=None, precision: str = "fp16",
                         domain_invariant: bool = False, batch_size: int = 250, num_workers: int = 0, rank: int = 0, world_size: int = 1,
                         progress: Optional[Callable] = None, **model_kwargs) -> dict:
    """results[dataset][checkpoint file] = {"top1", "top5", "num-samples"} for every checkpoint x dataset.

    `datasets` / `classnames`: name -> torch Dataset of (image, label) / list (or idx -> name dict) of class names.  One model is
    created once (`precision` defaults to the reference wrapper's fp16, xclip/open_clip/model.py:35) and every checkpoint is
    loaded into it.  With world_size > 1 call it from every rank of an initialised process group (one process per GPU, the
    current CUDA device set per rank, as the reference's classifier requires: xclip/zero_shot.py:17-18)."""
    from ..open_clip import create_model, get_tokenizer, load_checkpoint
    from ..open_clip.model import get_input_dtype
    from .open_clip import OpenCLIP
    model = create_model(model_name, precision=precision, device="cuda", **model_kwargs).eval()
    clip = OpenCLIP(model)
    tokenizer = tokenizer if tokenizer is not None else get_tokenizer(model_name)
    in_dtype = get_input_dtype(precision)
    results = {name: {} for name in datasets}
    for ckpt in ckpt_files:
        load_checkpoint(model, ckpt)
        for name, ds in datasets.items():
            results[name][str(ckpt)] = evaluate_model(clip, tokenizer, ds, classnames[name], domain_invariant, batch_size, num_workers, rank,
                                                      world_size, input_dtype=in_dtype)
        if progress is not None:
            progress(ckpt)
    return results
