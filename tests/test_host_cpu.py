"""CPU suite, part 2: host-side logic — the C-ABI library loads and exports every symbol of include/b200clip.h (no
compute calls without a GPU), the drop-in module mirrors the reference's state_dict / signatures / init / tokenizer,
the product path refuses to run without CUDA, and the multi-rank feature gather works under 2-process gloo."""
import inspect
import os
import re
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "b200clip.h"


# ------------------------------------------------------------------ C ABI ---------------------------
def _declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200clip_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from understanding_clip_ood_b200 import _lib
    lib = _lib.load()                                   # raises if the .so is missing: there is no fallback
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/b200clip.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes SIGNATURES out of sync with the header"
    assert lib.b200clip_version() >= 100
    assert lib.b200clip_launch_count() == 0 or lib.b200clip_launch_count() > 0   # callable without a device


def test_bad_arguments_are_reported_not_crashed():
    from understanding_clip_ood_b200 import _lib
    lib = _lib.load()
    rc = lib.b200clip_gemm(1, None, 0, None, 0, None, None, 0, None, 0, 0, 0, 0, 0, None, 0, 0, None)
    assert rc < 0 and b"null" in lib.b200clip_last_error()
    rc = lib.b200clip_zeroshot(0, None, None, None, None, None, 1, 1, 4, 1, 1, 1.0, None)
    assert rc < 0
    assert lib.b200clip_workspace_bytes(None, 1, 1) == -1


def test_workspace_bytes_formula():
    from understanding_clip_ood_b200 import _lib
    lib = _lib.load()
    import ctypes as C
    cfg = _lib.TowerCfg(dtype=1, width=768, layers=12, heads=12, mlp_width=3072, embed_dim=512, seq_len=50, quick_gelu=0,
                        image_size=224, patch_size=32, patch_kpad=3072, vocab_size=0)
    n = lib.b200clip_workspace_bytes(C.byref(cfg), 128, 50)
    rows = 128 * 50
    expect = rows * 768 * 2 * 2 + rows * 2304 * 2 + rows * 3072 * 2 + 128 * 768 * 2 + 128 * 4 + rows * 2 * 4   # + LN-fold row stats
    expect += 2 * rows * (768 // 64 + 2) * 8                 # + partial-sum slots of the two residual GEMMs (fused statistics)
    expect += lib.b200clip_gemm_workspace_bytes()             # + stream-K partial accumulators and flags of the CTA-pair GEMM
    assert lib.b200clip_gemm_workspace_bytes() == 74 * 2 * 64 * 128 * 16 + ((74 * 16 * 4 + 255) // 256) * 256   # 148 SMs assumed off-GPU
    assert expect <= n <= expect + 10 * 256


# ------------------------------------------------------------------ module surface ------------------
def test_no_cpu_fallback():
    from understanding_clip_ood_b200 import _lib, open_clip, ops
    m = open_clip.create_model("ViT-B-32", vision_cfg={"image_size": 64, "layers": 1, "width": 64, "patch_size": 32},
                               text_cfg={"context_length": 77, "vocab_size": 64, "width": 64, "heads": 1, "layers": 1},
                               embed_dim=64)
    with pytest.raises(_lib.B200ClipError):
        m.encode_image(torch.zeros(1, 3, 64, 64))
    with pytest.raises(_lib.B200ClipError):
        m.encode_text(torch.zeros(1, 77, dtype=torch.long))
    with pytest.raises(_lib.B200ClipError):
        ops.normalize(torch.zeros(2, 8))
    with pytest.raises(_lib.B200ClipError):
        open_clip.ClipLoss()(torch.zeros(4, 8), torch.zeros(4, 8), torch.tensor(1.0))


def test_state_dict_layout_vit_b_32():
    from understanding_clip_ood_b200 import open_clip
    m = open_clip.create_model("ViT-B-32", precision="bf16")
    sd = m.state_dict()
    assert len(sd) == 302
    assert sum(p.numel() for p in m.parameters()) == 151_277_313
    assert sd["visual.conv1.weight"].shape == (768, 3, 32, 32)
    assert sd["visual.transformer.resblocks.11.attn.in_proj_weight"].shape == (2304, 768)
    assert sd["transformer.resblocks.0.mlp.c_proj.weight"].shape == (512, 2048)
    assert sd["token_embedding.weight"].shape == (49408, 512) and sd["text_projection"].shape == (512, 512)
    assert "attn_mask" not in sd                                           # non-persistent buffer (model.py:248)
    n_lp = sum(v.dtype == torch.bfloat16 for v in sd.values())
    n_f32 = sum(v.dtype == torch.float32 for v in sd.values())
    assert (n_lp, n_f32) == (195, 107)                                     # convert_weights_to_lp split (probe)
    for k in ("visual.ln_pre.weight", "visual.class_embedding", "visual.positional_embedding", "positional_embedding",
              "token_embedding.weight", "logit_scale", "ln_final.bias"):
        assert sd[k].dtype == torch.float32
    assert abs(float(sd["logit_scale"]) - 2.6593) < 1e-3
    assert m.context_length == 77 and m.vocab_size == 49408 and m.visual.image_size == (224, 224)
    assert m.visual.preprocess_cfg["mean"][0] == pytest.approx(0.48145466)


def test_configs_and_errors():
    from understanding_clip_ood_b200 import open_clip
    assert {"ViT-B-32", "ViT-B-16", "ViT-L-14", "ViT-B-32-quickgelu"} <= set(open_clip.list_models())
    assert open_clip.get_model_config("ViT-L-14")["vision_cfg"]["width"] == 1024
    assert {"RN50", "RN101", "RN50x4", "RN50x16", "RN50x64", "RN50-quickgelu"} <= set(open_clip.list_models())
    with pytest.raises(RuntimeError):
        open_clip.create_model("convnext_base")
    with pytest.raises(RuntimeError):
        open_clip.create_model("ViT-B-32", precision="amp_fp8")
    tiny_kw = dict(vision_cfg={"image_size": 64, "layers": 1, "width": 64, "patch_size": 32},
                   text_cfg={"context_length": 77, "vocab_size": 64, "width": 64, "heads": 1, "layers": 1}, embed_dim=64)
    for prec, want in (("amp", torch.float16), ("amp_bf16", torch.bfloat16), ("amp_bfloat16", torch.bfloat16)):
        m = open_clip.create_model("ViT-B-32", precision=prec, **tiny_kw)      # fp32 master parameters, 16-bit kernels
        assert all(p.dtype == torch.float32 for p in m.parameters())
        assert m.visual._compute_dtype() == want and m._text_compute_dtype() == want
    m = open_clip.create_model("ViT-B-32", precision="fp32", **tiny_kw)
    assert m.visual._compute_dtype() == torch.float32
    with pytest.raises(RuntimeError):
        open_clip.create_model("ViT-B-32", pretrained="laion2b_s34b_b79k")
    m = open_clip.create_model("ViT-B-32-quickgelu", vision_cfg={"image_size": 64, "layers": 1, "width": 64, "patch_size": 32},
                               text_cfg={"context_length": 77, "vocab_size": 64, "width": 64, "heads": 1, "layers": 1},
                               embed_dim=64, output_dict=True)
    assert m.quick_gelu and m.visual.quick_gelu and m.output_dict
    m.set_grad_checkpointing()
    assert m.transformer.grad_checkpointing and m.visual.transformer.grad_checkpointing
    m.lock_image_tower()
    assert not any(p.requires_grad for p in m.visual.parameters())


def test_checkpoint_roundtrip(tmp_path):
    from understanding_clip_ood_b200 import open_clip
    kw = dict(vision_cfg={"image_size": 64, "layers": 1, "width": 64, "patch_size": 32},
              text_cfg={"context_length": 77, "vocab_size": 64, "width": 64, "heads": 1, "layers": 1}, embed_dim=64)
    a = open_clip.create_model("ViT-B-32", **kw)
    path = tmp_path / "epoch_1.pt"
    torch.save({"epoch": 1, "name": "x", "state_dict": {"module." + k: v for k, v in a.state_dict().items()}}, path)
    b = open_clip.create_model("ViT-B-32", pretrained=str(path), **kw)
    for k, v in a.state_dict().items():
        assert torch.equal(v, b.state_dict()[k])
    from understanding_clip_ood_b200.xclip.open_clip import OpenCLIP
    w, _, _ = OpenCLIP.from_pretrained("ViT-B-32", str(path), precision="fp32", **kw)
    assert torch.equal(w.clip.state_dict()["visual.proj"], a.state_dict()["visual.proj"])
    assert float(w.logit_scale) == pytest.approx(1 / 0.07, rel=1e-5)


# ------------------------------------------------------------------ pinned against the live reference
def _ref():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box)")
    return ref_loader.load()


def test_signatures_match_reference():
    ref_oc, ref_zs, ref_xo = _ref()
    from understanding_clip_ood_b200 import open_clip as mine
    from understanding_clip_ood_b200.xclip import open_clip as my_xo
    from understanding_clip_ood_b200.xclip import zero_shot as my_zs

    def params(fn):
        return [(p.name, p.default) for p in inspect.signature(fn).parameters.values() if p.kind != p.VAR_KEYWORD]

    assert params(mine.create_model) == params(ref_oc.create_model)
    assert params(mine.ClipLoss.__init__) == params(ref_oc.ClipLoss.__init__)
    assert params(mine.ClipLoss.forward) == params(ref_oc.ClipLoss.forward)
    assert params(mine.CLIP.encode_image) == params(ref_oc.CLIP.encode_image)
    assert params(mine.CLIP.encode_text) == params(ref_oc.CLIP.encode_text)
    assert params(mine.CLIP.forward) == params(ref_oc.CLIP.forward)
    assert [n for n, _ in params(mine.create_model_and_transforms)] == [n for n, _ in params(ref_oc.create_model_and_transforms)]
    for cls in ("ZeroShotClassifier", "OpenAIZeroShotClassifier"):
        assert [n for n, _ in params(getattr(my_zs, cls).__init__)] == [n for n, _ in params(getattr(ref_zs, cls).__init__)]
        for meth in ("predict", "predict_from_features", "variance_from_features"):
            assert params(getattr(getattr(my_zs, cls), meth)) == params(getattr(getattr(ref_zs, cls), meth))
    assert params(my_xo.OpenCLIP.from_pretrained) == params(ref_xo.OpenCLIP.from_pretrained)
    assert my_zs.OpenAIZeroShotClassifier.templates == ref_zs.OpenAIZeroShotClassifier.templates


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_seeded_init_equals_reference(precision):
    ref_oc, _, _ = _ref()
    from understanding_clip_ood_b200 import open_clip as mine
    kw = dict(embed_dim=64, vision_cfg={"image_size": 64, "layers": 2, "width": 128, "patch_size": 16},
              text_cfg={"context_length": 77, "vocab_size": 300, "width": 64, "heads": 1, "layers": 2})
    torch.manual_seed(123)
    r = ref_oc.create_model("ViT-B-32", precision=precision, **kw).state_dict()
    torch.manual_seed(123)
    m = mine.create_model("ViT-B-32", precision=precision, **kw).state_dict()
    assert list(r) == list(m)
    for k in r:
        assert r[k].dtype == m[k].dtype and torch.equal(r[k], m[k]), k


def test_seeded_init_equals_reference_full_vit_b_32():
    ref_oc, _, _ = _ref()
    from understanding_clip_ood_b200 import open_clip as mine
    torch.manual_seed(0)
    r = ref_oc.create_model("ViT-B-32").state_dict()
    torch.manual_seed(0)
    m = mine.create_model("ViT-B-32").state_dict()
    assert all(torch.equal(r[k], m[k]) for k in r) and list(r) == list(m)


@pytest.mark.parametrize("name,precision", [("RN50", "fp32"), ("RN50", "bf16"), ("RN50x4", "fp16")])
def test_seeded_init_equals_reference_modified_resnet(name, precision):
    """ModifiedResNet towers: same state_dict keys, order, dtypes (BatchNorm and the positional table stay fp32) and seed-0 values."""
    ref_oc, _, _ = _ref()
    from understanding_clip_ood_b200 import open_clip as mine
    torch.manual_seed(0)
    r = ref_oc.create_model(name, precision=precision).state_dict()
    torch.manual_seed(0)
    m = mine.create_model(name, precision=precision).state_dict()
    assert list(r) == list(m)
    for k in r:
        assert r[k].dtype == m[k].dtype and torch.equal(r[k], m[k]), k


def test_modified_resnet_host_logic():
    """Folding BatchNorm into the convolution operands (resnet.py:_fold) against the reference's conv -> bn on the CPU, the
    state_dict layout, and the no-CPU-fallback rule."""
    from understanding_clip_ood_b200 import open_clip
    from understanding_clip_ood_b200._lib import B200ClipError
    from understanding_clip_ood_b200.open_clip.resnet import _fold
    from oracle import clip_oracle as O
    kw = dict(embed_dim=128, vision_cfg={"image_size": 64, "layers": [1, 2, 1, 1], "width": 32, "patch_size": None},
              text_cfg={"context_length": 77, "vocab_size": 100, "width": 64, "heads": 1, "layers": 1})
    torch.manual_seed(0)
    m = open_clip.create_model("RN50", **kw).eval()
    keys = list(m.state_dict())
    assert "visual.layer2.0.downsample.1.running_var" in keys and "visual.attnpool.positional_embedding" in keys
    assert "visual.layer2.1.downsample.0.weight" not in keys and m.visual.attnpool.num_heads == 16
    assert all(float(b.bn3.weight.abs().max()) == 0 for b in m.visual.bottlenecks())      # modified_resnet.py:143-146
    O.randomize_batchnorm_(m.visual, 5)
    blk = m.visual.layer2[0]
    blk.bn2.running_mean.normal_(generator=torch.Generator().manual_seed(1))
    blk.bn2.running_var.uniform_(0.5, 2.0, generator=torch.Generator().manual_seed(2))
    x = torch.randn(2, blk.conv2.in_channels, 6, 6, generator=torch.Generator().manual_seed(3))
    want = blk.bn2(blk.conv2(x))                                                          # eval-mode BatchNorm
    w, b = _fold(blk.conv2, blk.bn2)
    cols = torch.nn.functional.unfold(x, 3, padding=1).reshape(2, x.shape[1], 9, 36).permute(0, 3, 2, 1).reshape(72, -1)   # K = (tap, c)
    got = (cols @ w.t() + b).reshape(2, 36, -1).permute(0, 2, 1).reshape(want.shape)
    assert float((got - want).abs().max()) < 1e-5
    w0, _ = _fold(m.visual.conv1, m.visual.bn1, kpad=32)
    assert w0.shape == (16, 32) and float(w0[:, 27:].abs().max()) == 0
    with pytest.raises(B200ClipError):
        m.encode_image(torch.zeros(1, 3, 64, 64))                                         # CPU tensors: no fallback


def test_tokenizer_equals_reference():
    ref_oc, ref_zs, _ = _ref()
    from understanding_clip_ood_b200 import open_clip as mine
    import json
    names = json.loads((ROOT / "understanding_clip_ood_b200" / "data" / "domainnet_classes.json").read_text())
    texts = [t.format(c) for c in names[::9] for t in ref_zs.OpenAIZeroShotClassifier.templates[::5]]
    texts += ["Hello, World!  it's  42 things &amp; more...", "naïve café ☕ 日本語", "x" * 400, ""]
    a, b = ref_oc.get_tokenizer("ViT-B-32")(texts), mine.get_tokenizer("ViT-B-32")(texts)
    assert torch.equal(a, b)
    tok = mine.get_tokenizer("ViT-B-32")
    assert tok.decode(tok.encode("a photo of a dog")).strip() == "a photo of a dog"


# ------------------------------------------------------------------ multi-rank host logic (gloo, CPU) -
def _gather_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from understanding_clip_ood_b200.open_clip.loss import gather_features, local_labels
    n, D = 4, 8
    g = torch.Generator().manual_seed(100 + rank)
    img = torch.randn(n, D, generator=g, requires_grad=True)
    txt = torch.randn(n, D, generator=g, requires_grad=True)
    all_img, all_txt = gather_features(img, txt, local_loss=True, gather_with_grad=True, rank=rank, world_size=world)
    # a loss whose gradient w.r.t. the gathered rows depends on the rank: backward must SUM over ranks (reduce-scatter)
    ((rank + 1) * (all_img.sum() + 2 * all_txt.sum())).backward()
    ng_img, ng_txt = gather_features(img, txt, local_loss=False, gather_with_grad=False, rank=rank, world_size=world)
    ret[rank] = {"img": img.detach(), "txt": txt.detach(), "all_img": all_img.detach(), "all_txt": all_txt.detach(),
                 "d_img": img.grad.clone(), "d_txt": txt.grad.clone(), "ng_requires": (ng_img.requires_grad, ng_txt.requires_grad),
                 "ng_img": ng_img.detach(), "labels": local_labels(n, rank, world, True)}
    dist.destroy_process_group()


def test_gather_features_two_rank_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_gather_worker, args=(2, 29541, ret), nprocs=2, join=True)
    r0, r1 = ret[0], ret[1]
    cat_img, cat_txt = torch.cat([r0["img"], r1["img"]]), torch.cat([r0["txt"], r1["txt"]])
    for r in (r0, r1):
        assert torch.equal(r["all_img"], cat_img) and torch.equal(r["all_txt"], cat_txt)     # rank-major order (loss.py:48-50)
        assert torch.equal(r["ng_img"], cat_img) and r["ng_requires"] == (True, True)        # local slice keeps its grad (:56-59)
        # d(all.sum)/d(local row) summed over ranks: (1 + 2) for img, 2 * (1 + 2) for txt
        assert torch.allclose(r["d_img"], torch.full_like(r["d_img"], 3.0))
        assert torch.allclose(r["d_txt"], torch.full_like(r["d_txt"], 6.0))
    assert r0["labels"].tolist() == [0, 1, 2, 3] and r1["labels"].tolist() == [4, 5, 6, 7]   # loss.py:92-94


# ------------------------------------------------------------------ bench.py contract (reference arm runs on CPU) ---
def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` = the reference's own code (baseline/_ref) — or, where that was not installed, the CPU oracle
    port — timed on the host cores; one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    env = dict(os.environ, B200CLIP_REF_SAMPLE="2")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "zero_shot_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    want_kind = "reference" if (ROOT / "baseline" / "_ref" / "open_clip" / "model.py").exists() else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_peer_exchange_ring_slot_bookkeeping():
    """Host logic of the peer-memory ClipLoss exchange that needs no GPU: a ring slot is held while an autograd ctx that saved
    it is alive, released by its backward (explicitly) or when the graph dies, and re-entering a held slot is refused."""
    import gc
    from understanding_clip_ood_b200.open_clip import peer

    class FakeState:
        def __init__(self):
            self.inflight = set()

    st = FakeState()
    tokens = [peer._SlotToken(st, s) for s in (1, 2, 3)]
    assert st.inflight == {1, 2, 3}
    tokens[0] = None                  # what `ctx.token = None` does at the end of backward
    gc.collect()
    assert st.inflight == {2, 3}
    del tokens                        # graphs that never ran backward
    gc.collect()
    assert st.inflight == set()
    assert peer.RING >= 2 and peer.MAX_WORLD == 16
    # B200CLIP_P2P=0 routes ClipLoss to the NCCL form
    old = os.environ.get("B200CLIP_P2P")
    try:
        os.environ["B200CLIP_P2P"] = "0"
        assert peer.enabled() is False and peer.get_exchange(4, 8, 0, 2, torch.device("cpu")) is None
        os.environ["B200CLIP_P2P"] = "1"
        assert peer.enabled() is True
    finally:
        if old is None:
            os.environ.pop("B200CLIP_P2P", None)
        else:
            os.environ["B200CLIP_P2P"] = old


def test_built_library_is_blackwell_native_sass():
    """The shipped kernels are sm_100a code on the Blackwell paths, not recompiled legacy kernels: SASS of the in-tree
    libb200clip.so (cuobjdump; mnemonic table of /opt/skills/guides/B200_PROFILING.md) shows tcgen05 MMAs (UTCHMMA) fed by TMA
    (UTMALDG) with TMEM loads / stores (LDTM / STTM) and TMA stores incl. the reduce-add form (UTMASTG / UTMAREDG) in the GEMM
    and the long-sequence attention kernel, and system-scope release / acquire accesses in the peer-memory exchange."""
    import collections
    import shutil
    import subprocess
    from understanding_clip_ood_b200 import _lib
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(tool).exists():
        pytest.skip("cuobjdump not available")
    _lib.load()
    sass = subprocess.run([tool, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True, timeout=600).stdout
    per = collections.defaultdict(collections.Counter)
    for part in re.split(r"\n\s*Function : ", sass)[1:]:
        name = part.split("\n", 1)[0]
        m = re.search(r"(gemm_pair_kernel|attention_tc_kernel|attention_short_kernel|p2p_allgather_kernel|p2p_reduce_finish_kernel)", name)
        if not m:
            continue
        for mn in ("UTCHMMA", "UTMALDG", "UTMASTG", "UTMAREDG", "LDTM", "STTM", "HMMA", "USETMAXREG", "STG.E.STRONG.SYS", "LDG.E.STRONG.SYS"):
            per[m.group(1)][mn] += len(re.findall(r"\b" + re.escape(mn), part))
    g = per["gemm_pair_kernel"]
    assert g["UTCHMMA"] > 0 and g["UTMALDG"] > 0 and g["UTMASTG"] > 0 and g["UTMAREDG"] > 0 and g["LDTM"] > 0 and g["HMMA"] == 0
    a = per["attention_tc_kernel"]
    assert a["UTCHMMA"] > 0 and a["UTMALDG"] > 0 and a["LDTM"] > 0 and a["STTM"] > 0 and a["USETMAXREG"] > 0 and a["HMMA"] == 0
    assert per["attention_short_kernel"]["UTMALDG"] > 0 and per["attention_short_kernel"]["UTMASTG"] > 0
    assert per["p2p_allgather_kernel"]["STG.E.STRONG.SYS"] > 0 and per["p2p_allgather_kernel"]["LDG.E.STRONG.SYS"] > 0
    assert per["p2p_reduce_finish_kernel"]["STG.E.STRONG.SYS"] > 0 and per["p2p_reduce_finish_kernel"]["LDG.E.STRONG.SYS"] > 0


def test_training_mode_takes_the_training_path_and_still_refuses_cpu_tensors():
    """In training mode with autograd on the towers go through their autograd nodes (open_clip/train.py); CPU tensors still get
    the loud no-fallback error, in training as in evaluation mode."""
    from understanding_clip_ood_b200 import _lib, open_clip
    m = open_clip.create_model("ViT-B-32", vision_cfg={"image_size": 64, "layers": 1, "width": 64, "patch_size": 32},
                               text_cfg={"context_length": 77, "vocab_size": 64, "width": 64, "heads": 1, "layers": 1}, embed_dim=32)
    for mode in (m.train, m.eval):
        mode()
        with pytest.raises(_lib.B200ClipError):
            m.encode_image(torch.zeros(1, 3, 64, 64))
        with pytest.raises(_lib.B200ClipError):
            m.encode_text(torch.zeros(1, 77, dtype=torch.long))


def test_ctypes_signatures_match_header_prototypes():
    """Parameter COUNT and coarse kinds (pointer / 64-bit integer / 32-bit integer / float) of every ctypes signature against
    the prototype in include/b200clip.h — a mismatch would not fail at load time, it would corrupt the call."""
    import ctypes as C
    from understanding_clip_ood_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    protos = dict(re.findall(r"\b(b200clip_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text))
    assert set(protos) == set(_lib.SIGNATURES)

    def kind_of_c(param: str) -> str:
        p = param.strip()
        if p in ("void", ""):
            return "none"
        if "*" in p:
            return "ptr"
        if re.search(r"\b(int64_t|uint64_t)\b", p):
            return "i64"
        if re.search(r"\bfloat\b", p):
            return "f32"
        if re.search(r"\bdouble\b", p):
            return "f64"
        if re.search(r"\b(int|int32_t|uint32_t)\b", p):
            return "i32"
        raise AssertionError(f"unrecognised parameter {p!r}")

    def kind_of_ctypes(t) -> str:
        if t in (C.c_void_p, C.c_char_p) or (isinstance(t, type) and issubclass(t, C._Pointer)):
            return "ptr"
        if t in (C.c_int64, C.c_uint64):
            return "i64"
        if t is C.c_float:
            return "f32"
        if t is C.c_double:
            return "f64"
        if t in (C.c_int, C.c_int32, C.c_uint32):
            return "i32"
        raise AssertionError(f"unrecognised ctypes type {t}")

    for name, params in protos.items():
        want = [k for k in (kind_of_c(p) for p in params.split(",")) if k != "none"]
        got = [kind_of_ctypes(t) for t in _lib.SIGNATURES[name][1]]
        assert got == want, f"{name}: header {want} vs ctypes {got}"


def test_openclip_from_pretrained_checkpoint_forms(tmp_path):
    """xclip OpenCLIP.from_pretrained (reference: xclip/open_clip/model.py:30-56): bare state dict, training checkpoint
    ({"state_dict": ...}) and DistributedDataParallel `module.` prefix all load; fp16 is the default precision; logit_scale is
    exp(param) clamped to [0, 100]."""
    import math
    from understanding_clip_ood_b200 import open_clip
    from understanding_clip_ood_b200.xclip.open_clip import OpenCLIP
    kw = dict(vision_cfg={"image_size": 64, "layers": 1, "width": 64, "patch_size": 32},
              text_cfg={"context_length": 77, "vocab_size": 64, "width": 64, "heads": 1, "layers": 1}, embed_dim=32)
    torch.manual_seed(3)
    src = open_clip.create_model("ViT-B-32", precision="fp32", **kw)
    sd = {k: v.clone() for k, v in src.state_dict().items()}
    forms = {"bare": sd, "wrapped": {"state_dict": sd, "epoch": 3}, "ddp": {"state_dict": {"module." + k: v for k, v in sd.items()}}}
    for name, blob in forms.items():
        path = tmp_path / f"{name}.pt"
        torch.save(blob, path)
        wrapper, pre_train, pre_val = OpenCLIP.from_pretrained("ViT-B-32", ckpt_path=str(path), precision="fp32", **kw)
        got = wrapper.clip.state_dict()
        assert set(got) == set(sd) and all(torch.equal(got[k], sd[k]) for k in sd), name
        assert callable(pre_train) and callable(pre_val)
    wrapper, _, _ = OpenCLIP.from_pretrained("ViT-B-32", **kw)
    assert wrapper.clip.visual.proj.dtype == torch.float16                       # default precision of the wrapper
    wrapper.clip.logit_scale.data.fill_(10.0)
    assert float(wrapper.logit_scale) == 100.0
    wrapper.clip.logit_scale.data.fill_(1.0)
    assert abs(float(wrapper.logit_scale) - math.e) < 1e-5
    assert wrapper.vocab_size == 64 and wrapper.uses_one_hot_encoding is False


def test_statistics_slot_count_fits_the_workspace_carve_up():
    """b200clip_gemm_stats_slots (2 per N tile of the tile shape the picker chooses; needs no device: the SM count defaults
    to 148) stays within the W/64 + 2 slots per row that the tower workspace reserves, for every tower width and batch."""
    from understanding_clip_ood_b200 import _lib
    lib = _lib.load()
    for W in (512, 768, 1024, 1280):
        for rows in (77, 6400, 25600, 51200, 65792, 806912):
            slots = lib.b200clip_gemm_stats_slots(rows, W)
            assert slots >= 2 and slots % 2 == 0 and slots <= W // 64 + 2, (W, rows, slots)
    assert lib.b200clip_gemm_stats_slots(51200, 768) == 6          # three 256-wide N tiles, two epilogue groups each
    assert lib.b200clip_gemm_stats_slots(0, 768) < 0               # invalid argument -> error code, no crash


def test_eval_driver_shards_batches_round_robin():
    """xclip/evaluate.py: rank r takes batches r, r + world, ...; together the ranks cover every batch exactly once."""
    from understanding_clip_ood_b200.xclip.evaluate import shard_batches
    batches = list(range(11))
    seen = []
    for r in range(4):
        mine = list(shard_batches(batches, r, 4))
        assert mine == batches[r::4]
        seen += mine
    assert sorted(seen) == batches
    assert list(shard_batches(batches)) == batches
    with pytest.raises(ValueError):
        list(shard_batches(batches, 4, 4))


@pytest.mark.parametrize("h,w,S", [(300, 400, 224), (500, 333, 224), (224, 224, 224), (231, 517, 224), (100, 150, 224), (768, 512, 336), (64, 96, 32)])
def test_resize_crop_tables_reproduce_pillow_bit_for_bit(h, w, S):
    """open_clip/gpu_transform.py restates how Pillow derives its fixed-point bicubic tables; the two integer passes evaluated in
    numpy must equal torchvision Resize(S, BICUBIC) + CenterCrop(S) on the PIL image — the eval preprocessing the reference builds
    (deps/open_clip/src/open_clip/transform.py:372-392) — bit for bit, for down- and up-scaling, both orientations."""
    from PIL import Image
    from torchvision import transforms as T
    from torchvision.transforms import InterpolationMode
    from understanding_clip_ood_b200.open_clip import gpu_transform as G
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    img[: h // 2] = (np.linspace(0, 255, w)[None, :, None] * np.ones((h // 2, 1, 3))).astype(np.uint8)     # a smooth half as well
    ref = np.asarray(T.Compose([T.Resize(S, interpolation=InterpolationMode.BICUBIC), T.CenterCrop(S)])(Image.fromarray(img))).transpose(2, 0, 1)
    assert np.array_equal(G.emulate_numpy(img, S), ref)
    assert G.resized_size(h, w, S) == tuple(T.Resize(S)(Image.fromarray(img)).size[::-1])


def test_parameter_snapshot_tracks_the_module_tree_exactly():
    """open_clip/model.py:_ParamSnapshot — the cached parameter list every encode call uses instead of walking the module tree
    (0.4 ms per walk for a 150-tensor tower) must equal a fresh `parameters()` traversal after ANY structural change: a replaced
    Parameter, a replaced sub-module, an added parameter; dtype / device casts and optimizer steps keep the Parameter objects and
    are the engine signature's business."""
    import torch
    from understanding_clip_ood_b200 import open_clip
    from understanding_clip_ood_b200.open_clip import model as M
    m = open_clip.create_model("ViT-B-32", precision="fp32", device="cpu",
                               embed_dim=32, vision_cfg={"image_size": 32, "layers": 3, "width": 128, "patch_size": 16},
                               text_cfg={"context_length": 8, "vocab_size": 50, "width": 128, "heads": 2, "layers": 2})
    v = m.visual

    def same():
        snap = M._snapshot(v)
        return [id(p) for p in snap.params] == [id(p) for p in v.parameters()] and snap.names == [n for n, _ in v.named_parameters()]

    assert same()
    first = M._snapshot(v).params
    assert M._snapshot(v).params is first                        # nothing changed: no re-collection
    blk = v.transformer.resblocks[1]
    blk.ln_1.weight = torch.nn.Parameter(torch.ones_like(blk.ln_1.weight))
    assert same() and M._snapshot(v).params is not first         # a replaced Parameter object
    v.transformer.resblocks[0] = v.transformer.resblocks[2]      # a replaced (here: shared) sub-module
    assert same()
    v.extra = torch.nn.Parameter(torch.zeros(3))                 # an added parameter
    assert same()
    v.half()
    assert same()                                                # casts swap .data, not the Parameter objects
    names_t = [n for n, _ in m._text_named_parameters()]
    assert names_t[:5] == ["token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection"]
    assert names_t[5:] == ["transformer." + n for n, _ in m.transformer.named_parameters()]
    sig, static_sig = M._Engine.signatures(list(v.parameters()), True, torch.float16)
    with torch.no_grad():
        v.proj.add_(1)
    sig2, static2 = M._Engine.signatures(list(v.parameters()), True, torch.float16)
    assert static2 == static_sig and sig2 != sig                 # an in-place update changes only the version part
