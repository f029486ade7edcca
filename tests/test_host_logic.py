"""CPU tests (no GPU): the C-ABI library loads and exports every symbol include/b200mart.h declares, the parameter
layout matches the reference state_dict contract, the product path fails loudly without CUDA, and the host-side
data-parallel logic (world size 2, gloo) averages shard gradients into the global-batch gradient."""
import os
import re
import socket

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mkg_analogy_b200 import _lib
    with open(os.path.join(ROOT, "include", "b200mart.h")) as fh:
        header = fh.read()
    declared = set(re.findall(r"\b(b200mart_[a-z0-9_]+)\s*\(", header))
    declared -= {"b200mart_handle"}
    assert len(declared) >= 20
    l = _lib.lib()                                  # resolves every SYMBOLS entry or raises
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert getattr(l, name) is not None
    assert l.b200mart_version() == 1
    assert l.b200mart_launch_count() == 0          # nothing launched: no compute without a GPU


def test_ctypes_structs_match_header_field_order():
    """Field names of each ctypes mirror appear in the header struct in the same order."""
    from mkg_analogy_b200 import _lib
    with open(os.path.join(ROOT, "include", "b200mart.h")) as fh:
        header = fh.read()
    for cname, cls in (("b200mart_gemm_desc", _lib.GemmDesc), ("b200mart_model_desc", _lib.ModelDesc),
                       ("b200mart_buffers", _lib.Buffers), ("b200mart_batch", _lib.Batch), ("b200mart_head", _lib.Head),
                       ("b200mart_attn_desc", _lib.AttnDesc), ("b200mart_stack_desc", _lib.StackDesc), ("b200mart_stack_io", _lib.StackIO)):
        body = re.search(r"typedef struct " + cname + r" \{(.*?)\} " + cname + ";", header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            parts = decl.split(",")
            names.append(re.findall(r"[A-Za-z_0-9]+", parts[0])[-1])
            names += [re.findall(r"[A-Za-z_0-9]+", p)[-1] for p in parts[1:]]
        assert names == [f[0] for f in cls._fields_], (cname, names, [f[0] for f in cls._fields_])


def test_param_layout_matches_reference_contract():
    from mkg_analogy_b200.config import BASE, LARGE, TINY, param_shapes
    from mkg_analogy_b200.engine import query_layout
    for cfg in (TINY, BASE, LARGE):
        infos, total = query_layout(cfg)
        ref = param_shapes(cfg)
        assert {p.name: tuple(p.shape) for p in infos} == {k: tuple(v) for k, v in ref.items()}
        # q/k/v projections of a layer are adjacent (one [3H, H] operand), every tensor starts 128-byte aligned in bf16
        by = {p.name: p for p in infos}
        H = cfg.hidden
        for pre, names in (("unimo.encoder.text_layer.0.attention.self.", ("query", "key", "value")),
                           ("unimo.encoder.vision_layers.0.self_attn.", ("q_proj", "k_proj", "v_proj"))):
            offs = [by[pre + n + ".weight"].offset for n in names]
            assert offs[1] - offs[0] == H * H and offs[2] - offs[1] == H * H
            boffs = [by[pre + n + ".bias"].offset for n in names]
            assert boffs[1] - boffs[0] == H and boffs[2] - boffs[1] == H
        assert all(p.offset % 64 == 0 for p in infos) and total % 64 == 0
        # reference weight-decay rule (lit_models/transformer.py:225-230) and never-trained parameters
        for p in infos:
            assert p.decay == (not ("bias" in p.name or "LayerNorm.weight" in p.name)), p.name
        unused = {p.name for p in infos if p.unused_ft}
        assert "unimo.text_pooler.dense.weight" in unused and "unimo.vision_post_layernorm.bias" in unused
        assert f"unimo.encoder.text_layer.{cfg.fuse_from - 1}.intermediate.fusion_dense.weight" in unused
        assert f"unimo.encoder.text_layer.{cfg.fuse_from}.intermediate.fusion_dense.weight" not in unused


def test_no_cpu_fallback():
    """Without CUDA the module is parameter storage only: any compute raises (never routes through the oracle)."""
    from mkg_analogy_b200._lib import B200MartError
    from mkg_analogy_b200.config import TINY
    from mkg_analogy_b200.modeling import MKGformerKGC
    m = MKGformerKGC(TINY, device="cpu")
    with pytest.raises(B200MartError):
        m(input_ids=torch.zeros(1, 16, dtype=torch.long), attention_mask=torch.ones(1, 16, dtype=torch.long),
          pixel_values=torch.zeros(1, 2, 3, 64, 64))
    import mkg_analogy_b200
    src = ""
    for f in os.listdir(os.path.dirname(mkg_analogy_b200.__file__)):
        if f.endswith(".py"):
            with open(os.path.join(os.path.dirname(mkg_analogy_b200.__file__), f)) as fh:
                src += fh.read()
    assert "import oracle" not in src and "from oracle" not in src


def test_module_mirrors_reference_interface():
    """state_dict names, tied embeddings, resize semantics, argparse hook."""
    import argparse
    from mkg_analogy_b200.config import TINY, param_shapes
    from mkg_analogy_b200.modeling import MKGformerKGC
    m = MKGformerKGC(TINY, device="cpu")
    sd = m.state_dict()
    for k, shp in param_shapes(TINY).items():
        assert tuple(sd[k].shape) == tuple(shp)
    assert {"cls.predictions.decoder.weight", "cls.predictions.decoder.bias", "unimo.text_embeddings.position_ids",
            "unimo.vision_embeddings.position_ids"} <= set(sd)
    assert len(list(m.named_parameters())) == len(param_shapes(TINY))          # aliases are not double counted
    old = m.get_input_embeddings().weight.detach().clone()
    oldb = m.cls.predictions.bias.detach().clone()
    with torch.no_grad():
        m.cls.predictions.bias.normal_()
        oldb = m.cls.predictions.bias.detach().clone()
    emb = m.resize_token_embeddings(TINY.vocab + 5)
    assert emb.weight.shape[0] == TINY.vocab + 5 and m.get_output_embeddings().weight is emb.weight
    assert torch.equal(emb.weight[:TINY.vocab], old)
    assert torch.equal(m.cls.predictions.bias[:TINY.vocab], oldb) and m.cls.predictions.bias[TINY.vocab:].abs().max() == 0
    with torch.no_grad():                                                      # what _init_relation_word does (:46-54)
        emb.weight[TINY.vocab + 4] = emb.weight[[3, 4, 5]].mean(0)
    assert torch.equal(m.get_input_embeddings().weight, m.get_output_embeddings().weight)
    m.load_state_dict(m.state_dict())
    p = argparse.ArgumentParser()
    MKGformerKGC.add_to_argparse(p)
    assert p.parse_args(["--pretrain", "1"]).pretrain == 1


def test_lr_schedule_matches_oracle_and_hf():
    """configure_optimizers (transformer.py:233-240) = HF get_linear_schedule_with_warmup with a FLOAT warm-up length."""
    from transformers.optimization import get_linear_schedule_with_warmup
    from mkg_analogy_b200.trainer import linear_warmup_lr
    from oracle.mkgformer_oracle import linear_warmup_lr as ref
    for total in (10, 100, 997):
        p = torch.nn.Parameter(torch.zeros(1))
        opt = torch.optim.AdamW([p], lr=5e-5)
        sch = get_linear_schedule_with_warmup(opt, num_warmup_steps=total * 0.1, num_training_steps=total)
        for step in range(0, total + 2):
            a, b, c = linear_warmup_lr(step, total, 5e-5, 0.1), ref(step, total, 5e-5, 0.1), sch.get_last_lr()[0]
            assert abs(a - b) <= 1e-15 and abs(a - c) <= 1e-12, (total, step, a, b, c)
            opt.step()
            sch.step()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    from mkg_analogy_b200.config import MKGformerConfig, init_params
    from mkg_analogy_b200.data import TINY_VOCAB, make_finetune_batch
    from mkg_analogy_b200.trainer import allreduce_mean_, bucket_bounds
    from oracle import mkgformer_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = MKGformerConfig(hidden=128, layers=9, heads=2, intermediate=256, vocab=TINY_VOCAB.vocab, max_position=64,
                          image_size=64, patch=32)
    P = {k: v.requires_grad_(True) for k, v in init_params(cfg, 5).items()}
    full = make_finetune_batch(4, 24, cfg.image_size, TINY_VOCAB, seed=77, pad=False)     # global batch
    shard = {k: v[rank * 2:(rank + 1) * 2] for k, v in full.items()}                     # this rank's equal-sized shard
    loss, _, _ = O.finetune_loss(P, cfg, shard, TINY_VOCAB.analogy_entity_ids())
    loss.backward()
    names = [n for n in P if P[n].grad is not None]
    flat = torch.cat([P[n].grad.reshape(-1) for n in names])
    assert bucket_bounds(flat.numel(), 5)[-1][1] == flat.numel()
    allreduce_mean_(flat, None, buckets=5)
    if rank == 0:
        Pf = {k: v.detach().clone().requires_grad_(True) for k, v in P.items()}
        lf, _, _ = O.finetune_loss(Pf, cfg, full, TINY_VOCAB.analogy_entity_ids())
        lf.backward()
        ref = torch.cat([Pf[n].grad.reshape(-1) for n in names])
        out.put(((flat - ref).abs().max().item(), ref.abs().max().item()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_mean_of_shards_is_global_gradient_gloo():
    """World size 2 on CPU (gloo): bucketed mean all-reduce of per-shard gradients == gradient of the global batch
    (both loss terms are batch means)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale = out.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err <= 1e-5 * scale + 1e-7, (err, scale)


def test_layer_extents_partition_the_encoder():
    """b200mart_layer_extent: contiguous, ordered, disjoint; exactly the encoder-layer parameters fall inside."""
    import ctypes as C
    from mkg_analogy_b200 import _lib
    from mkg_analogy_b200.config import BASE
    from mkg_analogy_b200.engine import _model_desc, _read_layout
    l = _lib.lib()
    h = _lib.vp()
    d = _model_desc(BASE)
    _lib.check(l.b200mart_create(C.byref(d), C.byref(h)))
    infos, total = _read_layout(l, h)
    ext = []
    for i in range(BASE.layers):
        b, e = _lib.i64(), _lib.i64()
        _lib.check(l.b200mart_layer_extent(h, i, C.byref(b), C.byref(e)))
        ext.append((b.value, e.value))
    l.b200mart_destroy(h)
    assert all(ext[i][1] == ext[i + 1][0] for i in range(len(ext) - 1)) and 0 < ext[0][0] and ext[-1][1] < total
    for p in infos:
        inside = [i for i, (b, e) in enumerate(ext) if b <= p.offset < e]
        if ".encoder." in p.name:
            layer = int(p.name.split(".")[3])
            assert inside == [layer], p.name
        else:
            assert inside == [], p.name


@pytest.mark.parametrize("family", ["vilt", "flava", "visual_bert"])
def test_from_pretrained_local_checkpoint(tmp_path, family):
    """`model_class.from_pretrained(path[, config=config])` (MarT/main.py:119-125) on a local HF-layout directory:
    config.json + safetensors / .bin weights load by name, absent tensors are initialised HF-style, the decoder stays tied."""
    import os
    from safetensors.torch import save_file
    GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    c = torch.load(os.path.join(GOLD, f"{family}_tiny.pt"), weights_only=False)
    if family == "vilt":
        from transformers import ViltConfig as Cfg
        from mkg_analogy_b200.modeling_vilt import ViltKGC as M
        head = "mlm_score."
    elif family == "visual_bert":
        from transformers import VisualBertConfig as Cfg
        from mkg_analogy_b200.modeling_visual_bert import VisualBertKGC as M
        head = "cls."
    else:
        from transformers import FlavaConfig as Cfg
        from mkg_analogy_b200.modeling_flava import FlavaKGC as M
        head = "cls."
    Cfg(**c["cfg"]).save_pretrained(tmp_path)
    body = {k: v.contiguous() for k, v in c["params"].items() if not k.startswith(head) and "position_ids" not in k}
    if family == "vilt":
        save_file(body, os.path.join(tmp_path, "model.safetensors"))
    else:
        torch.save(body, os.path.join(tmp_path, "pytorch_model.bin"))
    m = M.from_pretrained(str(tmp_path), device="cpu")
    sd = m.state_dict()
    for k, v in body.items():
        assert torch.equal(sd[k], v), k
    assert all(k.startswith(head) or "position_ids" in k for k in m.loading_info["missing_keys"]), m.loading_info
    assert m.get_output_embeddings().weight is m.get_input_embeddings().weight
    tr = head + ("predictions.transform.dense." if family == "visual_bert" else "transform.dense.")
    dense = sd[tr + "weight"]
    assert 0.01 < dense.std().item() < 0.03 and sd[tr + "bias"].abs().max().item() == 0.0
    m2 = M.from_pretrained(str(tmp_path), config=Cfg(**c["cfg"]), device="cpu")     # explicit config, as main.py:123 passes it
    assert torch.equal(m2.state_dict()[next(iter(body))], body[next(iter(body))])


def test_step_block_packing_and_bucket_choice():
    """Host side of the graph-replay mode: the 16-byte step block {uint32 salt, float lr, float 1 - beta1^t,
    float 1 - beta2^t} as include/b200mart.h lays it out, the salt = the kernels' hash32 finaliser (the oracle restates it), buckets."""
    import struct
    from mkg_analogy_b200.trainer import pack_step_block, pick_bucket, step_salt

    def h32(x):
        x ^= x >> 16
        x = (x * 0x7feb352d) & 0xFFFFFFFF
        x ^= x >> 15
        x = (x * 0x846ca68b) & 0xFFFFFFFF
        return x ^ (x >> 16)
    salts = [step_salt(1234567, t) for t in range(1, 200)]
    assert salts[0] == h32((1234567 + 0x9E3779B9) & 0xFFFFFFFF) and len(set(salts)) == len(salts) and all(0 <= x < 2 ** 32 for x in salts)
    blk = pack_step_block(torch.zeros(4, dtype=torch.int32), 0xDEADBEEF, 3e-5, 7)
    salt, lr, bc1, bc2 = struct.unpack("<Ifff", blk.numpy().tobytes())
    assert salt == 0xDEADBEEF and abs(lr - 3e-5) < 1e-12 and abs(bc1 - (1 - 0.9 ** 7)) < 1e-7 and abs(bc2 - (1 - 0.999 ** 7)) < 1e-9
    assert [pick_bucket(L, (32, 64, 96, 128)) for L in (1, 32, 33, 64, 100, 128)] == [32, 32, 64, 64, 128, 128]
    with pytest.raises(ValueError):
        pick_bucket(129, (32, 64, 96, 128))
