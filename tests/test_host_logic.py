"""CPU tests: host-side logic, state_dict compatibility, packing layouts, C-ABI symbol export.
No GPU compute is invoked here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import mintime_b200
from mintime_b200 import _lib, spec, synth, weights
from mintime_b200.efficientnet import EfficientNet
from mintime_b200.size_invariant_timesformer import SizeInvariantTimeSformer

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "mintime_b200.h")).read()
    declared = set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)            # must exist: built by __graft_entry__.build()
    for name in declared:
        assert hasattr(lib, name), name
    lib.mt_abi_version.restype = ctypes.c_int
    assert lib.mt_abi_version() == 5
    lib.mt_last_error.restype = ctypes.c_char_p
    assert lib.mt_last_error() == b""


def test_argument_validation_without_gpu():
    """invalid shapes are rejected on the host before any CUDA call (status < 0 + message)."""
    lib = _lib.load()
    rc = lib.mt_layernorm_fwd(1, 8, 8, 8, 8, 4, 100, None)      # dim not a multiple of 128
    assert rc == -1 and b"multiple of 128" in lib.mt_last_error()
    rc = lib.mt_pointwise_fwd(1, 8, 8, None, None, 0, None, 0, 8, 4, 12, 16, None)   # N % 8 != 0
    assert rc == -1 and b"multiples of 8" in lib.mt_last_error()
    rc = lib.mt_divided_attn_fwd(1, 8, 8, 8, 0, 8, None, 1, 16, 49, 8, 32, None, 0, None)     # dim_head != 64
    assert rc == -1 and b"dim_head" in lib.mt_last_error()
    assert lib.mt_effnet_b0_workspace_bytes(0, 1) == 0
    bf16_ws, fp32_ws = lib.mt_effnet_b0_workspace_bytes(4, 1), lib.mt_effnet_b0_workspace_bytes(4, 0)
    assert 0 < bf16_ws < fp32_ws <= 2.1 * bf16_ws          # fp32 activations are twice as wide


def test_effnet_state_dict_matches_reference_names():
    sd = synth.make_effnet_state_dict(1)
    m = EfficientNet.from_name("efficientnet-b0")
    own = m.state_dict()
    assert len(own) == 360 and set(own) == set(sd)
    for k in own:
        assert own[k].shape == sd[k].shape, k
    m.load_state_dict(sd, strict=True)
    # names parsed by the reference's partial-unfreeze logic (train.py:159-167)
    for name, _ in m.named_parameters():
        if name.startswith("_blocks."):
            assert 0 <= int(name.split(".")[1]) < 16
    with pytest.raises(ValueError):
        EfficientNet.from_name("efficientnet-b9")


def test_load_matching_state_dict_semantics():
    sd = synth.make_effnet_state_dict(2)
    m = EfficientNet.from_name("efficientnet-b0")
    renamed = {"efficient_net." + k: v for k, v in sd.items()}
    renamed["some.unknown.key"] = torch.zeros(3)
    m.load_matching_state_dict(renamed)               # model.py:368-378
    assert torch.equal(m.state_dict()["_blocks.3._project_conv.weight"], sd["_blocks.3._project_conv.weight"])


def test_module_prefixed_checkpoints_load_like_predict_py():
    """predict.py:375-388 wraps both modules in nn.DataParallel and loads checkpoints whose keys carry 'module.'
    (train.py:461-464 saves the wrapped modules).  Both routes work: into the wrapper, and into the bare module."""
    cfg = spec.default_tsf_config(num_frames=8)
    sd = synth.make_tsf_state_dict(cfg, 5)
    pref = {"module." + k: v for k, v in sd.items()}
    bare = SizeInvariantTimeSformer(config=cfg)
    res = bare.load_state_dict(pref)                            # strict
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(bare.state_dict()["layers.3.1.fn.to_qkv.weight"], sd["layers.3.1.fn.to_qkv.weight"])
    wrapped = torch.nn.DataParallel(SizeInvariantTimeSformer(config=cfg))
    wrapped.load_state_dict(pref)
    assert torch.equal(wrapped.module.state_dict()["to_out.1.weight"], sd["to_out.1.weight"])
    assert set(wrapped.state_dict()) == set(pref)
    esd = synth.make_effnet_state_dict(6)
    ext = EfficientNet.from_name("efficientnet-b0")
    ext.load_state_dict({"module." + k: v for k, v in esd.items()})
    assert torch.equal(ext.state_dict()["_blocks.7._bn1.running_var"], esd["_blocks.7._bn1.running_var"])
    with pytest.raises(RuntimeError):                           # strict loading still rejects foreign keys
        bare.load_state_dict({**pref, "module.nope": torch.zeros(1)})


def test_tsf_state_dict_matches_reference_names():
    for f in (8, 16):
        cfg = spec.default_tsf_config(num_frames=f)
        sd = synth.make_tsf_state_dict(cfg, 3)
        m = SizeInvariantTimeSformer(config=cfg, require_attention=True)
        own = m.state_dict()
        assert set(own) == set(sd)
        for k in own:
            assert own[k].shape == sd[k].shape, k
        m.load_state_dict(sd, strict=True)
        assert sum(p.numel() for p in m.parameters()) == (68_894_721 if f == 16 else None) or f == 8
        assert m.no_weight_decay() == {"pos_emb", "cls_token", "size_emb"}


def test_tsf_config_quirks():
    cfg = spec.default_tsf_config()
    cfg["model"]["shift-tokens"] = True
    with pytest.raises(NotImplementedError):
        SizeInvariantTimeSformer(config=cfg)
    cfg = spec.default_tsf_config()
    m = SizeInvariantTimeSformer(config=cfg)
    with pytest.raises(ValueError):       # f must equal config num-frames (reference :252)
        m(torch.zeros(1, 8, 1280, 7, 7), mask=torch.ones(1, 8, dtype=torch.bool),
          identities_mask=torch.ones(1, 8, 8, dtype=torch.bool), size_embedding=torch.zeros(1, 8, dtype=torch.int32),
          positions=torch.zeros(1, 393, dtype=torch.int64))


def test_no_cpu_fallback():
    """without a CUDA device the product path raises instead of computing on the CPU."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = EfficientNet.from_name("efficientnet-b0").eval()
    with pytest.raises(_lib.MintimeError):
        m(torch.zeros(1, 3, 224, 224))
    cfg = spec.default_tsf_config()
    t = SizeInvariantTimeSformer(config=cfg)
    with pytest.raises(_lib.MintimeError):
        t(torch.zeros(1, 16, 1280, 7, 7), mask=torch.ones(1, 16, dtype=torch.bool),
          identities_mask=torch.ones(1, 16, 16, dtype=torch.bool),
          size_embedding=torch.zeros(1, 16, dtype=torch.int32), positions=torch.zeros(1, 785, dtype=torch.int64))
    t.train()                                          # the training step (autograd node) has no CPU route either
    assert torch.is_grad_enabled() and all(p.requires_grad for p in t.parameters())
    with pytest.raises(_lib.MintimeError):
        t(torch.zeros(1, 16, 1280, 7, 7), mask=torch.ones(1, 16, dtype=torch.bool),
          identities_mask=torch.ones(1, 16, 16, dtype=torch.bool),
          size_embedding=torch.zeros(1, 16, dtype=torch.int32), positions=torch.zeros(1, 785, dtype=torch.int64))
    from mintime_b200 import ops
    with pytest.raises(_lib.MintimeError):
        ops.layernorm_bwd_(torch.zeros(8, 512), torch.zeros(8, 512), torch.ones(512), torch.zeros(8, 512).bfloat16())
    with pytest.raises(_lib.MintimeError):
        ops.grad_prep(torch.zeros(64, 64), want_t=True)
    e = EfficientNet.from_name("efficientnet-b0")      # train mode (batch-stat BN, drop-connect, backward): no CPU route either
    assert e.training
    with pytest.raises(_lib.MintimeError):
        e(torch.zeros(1, 3, 224, 224))


def test_block_table_in_csrc_matches_spec():
    src = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc", "effnet.cu")).read()
    body = src[src.index("kBlocks[16] = {"):]
    body = body[:body.index("};")]
    rows = re.findall(r"\{(\d+), (\d+), (\d+), (\d+), (\d+), (\d+)\}", body)
    assert len(rows) == 16
    for r, s in zip(rows, spec.B0_BLOCKS):
        assert tuple(map(int, r)) == (s.kernel, s.stride, s.expand, s.cin, s.cout, s.hw_in)
    # TF-SAME low-side padding per block (last column)
    assert [b.pad_lo for b in spec.B0_BLOCKS] == [1, 0, 1, 1, 2, 0, 1, 1, 2, 2, 2, 1, 2, 2, 2, 1]


def test_bn_fold_and_layouts():
    sd = synth.make_effnet_state_dict(5)
    pk = weights.pack_effnet(sd, "fp32", "cpu")
    # tensors are kept in pack order: stem w, stem shift, ...
    stem_w, stem_shift = pk.keep[0], pk.keep[1]
    scale = sd["_bn0.weight"] / torch.sqrt(sd["_bn0.running_var"] + spec.BN_EPS)
    ref = (sd["_conv_stem.weight"] * scale[:, None, None, None]).permute(2, 3, 1, 0).reshape(27, 32)
    assert torch.allclose(stem_w, ref, atol=1e-7)
    assert torch.allclose(stem_shift, sd["_bn0.bias"] - sd["_bn0.running_mean"] * scale, atol=1e-6)
    x = torch.randn(2, 3, 8, 8)
    y_ref = torch.nn.functional.batch_norm(
        torch.nn.functional.conv2d(x, sd["_conv_stem.weight"]), sd["_bn0.running_mean"], sd["_bn0.running_var"],
        sd["_bn0.weight"], sd["_bn0.bias"], False, 0.0, spec.BN_EPS)
    w = stem_w.reshape(3, 3, 3, 32).permute(3, 2, 0, 1)
    y = torch.nn.functional.conv2d(x, w) + stem_shift[None, :, None, None]
    assert torch.allclose(y, y_ref, atol=1e-4, rtol=1e-4)


def test_geglu_interleave_roundtrip():
    t = torch.arange(256 * 3, dtype=torch.float32).reshape(256, 3)
    p = weights.geglu_interleave(t)
    h = 128
    for r in range(256):
        blk, within = divmod(r, 64)
        src = blk * 32 + within if within < 32 else h + blk * 32 + (within - 32)
        assert torch.equal(p[r], t[src])


def test_pack_tsf_layout():
    cfg = spec.default_tsf_config(num_frames=8)
    sd = synth.make_tsf_state_dict(cfg, 7)
    sd = {"module." + k: v for k, v in sd.items()}      # DataParallel checkpoint (predict.py:378-388)
    pk = weights.pack_tsf(sd, cfg, "bf16", "cpu")
    assert pk.struct.w_patch and pk.struct.size_emb and pk.struct.ff[8].w2 and not pk.struct.ff[9].w2
    wq = [t for t in pk.keep if t.shape == (1536, 512)][0]
    ref = sd["module.layers.0.0.fn.to_qkv.weight"].clone()
    ref[:512] *= 0.125
    assert torch.equal(wq, ref.to(torch.bfloat16))


def test_synthetic_clip_metadata_contract():
    for ids, f in ((1, 16), (2, 16), (3, 16), (4, 16), (2, 8)):
        meta = synth.make_batch_meta(3, f, [ids], seed=5)
        m, im, se, pos = meta["mask"], meta["identities_mask"], meta["size_embedding"], meta["positions"]
        assert m.shape == (3, f) and im.shape == (3, f, f) and se.shape == (3, f) and pos.shape == (3, 1 + f * 49)
        assert m.dtype == torch.bool and im.dtype == torch.bool and se.dtype == torch.int32 and pos.dtype == torch.int64
        assert (pos[:, 0] == 0).all() and pos.max() <= f * 49 and pos[:, 1:].min() >= 1
        assert sum(synth.identity_slots(f, ids)) == f
        assert torch.equal(im, im.transpose(1, 2))                      # block diagonal
        assert ((se == 0) == (~m)).all()                                # padded slots carry size 0
        assert im.diagonal(dim1=1, dim2=2).all()
    assert synth.identity_slots(16, 3) == [5, 5, 6] and synth.identity_slots(16, 4) == [5, 5, 2, 4]


def test_clip_meta_oracle_known_answer():
    """The case tests/golden/clip_meta_ref.json holds as 'two_ids_padded_f8' (outputs of the executed reference), with
    2 patches per frame: f=8, identity A owns 4 slots with 4 faces, identity B owns 4 slots with 2 faces; the 2 padded
    slots repeat the CLIP's largest frame so far (12, not B's own 9), get size 0, and mask 0 only in predict.py."""
    from oracle.clip_meta_oracle import clip_meta
    ids = [(4, [(3, 0), (5, 5), (9, 6), (12, 100)]), (4, [(5, 17), (9, 50)])]
    se, mask, idm, pos = clip_meta(ids, 8, num_patches=2, source="predict")
    assert se.tolist() == [1, 1, 2, 20, 4, 10, 0, 0]
    assert mask.tolist() == [True] * 6 + [False] * 2
    assert idm[:4, :4].all() and idm[4:, 4:].all() and not idm[:4, 4:].any() and not idm[4:, :4].any()
    # distinct frames sorted: 3, 5, 9, 12 -> ranks 1..4; slots: 3,5,9,12 | 5,9,12,12
    ranks = [1, 2, 3, 4, 2, 3, 4, 4]
    want = [0] + [t for r in ranks for t in ((r - 1) * 2 + 1, (r - 1) * 2 + 2)]
    assert pos.tolist() == want
    # DeepFakesDataset as executed: every slot valid, whatever enable_identity_attention says (:283 vs :276)
    for ia in (True, False):
        se2, mask2, _, pos2 = clip_meta(ids, 8, num_patches=2, enable_identity_attention=ia, source="dataset")
        assert mask2.all() and se2.tolist() == se.tolist() and pos2.tolist() == want


def test_clip_meta_oracle_agrees_with_synthetic_generator():
    """synth.make_clip_meta (the bench / fixture input generator) and the line-by-line oracle describe the same
    assembly: rebuild the generator's clips through the oracle."""
    import numpy as np
    from oracle.clip_meta_oracle import clip_meta
    from mintime_b200 import synth
    for n_id in (1, 2, 3, 4):
        rng = np.random.default_rng(5 + n_id)
        se, mask, idm, pos = synth.make_clip_meta(16, n_id, rng, pad_tail=True)
        slots = synth.identity_slots(16, n_id)
        ids, start = [], 0
        ranks = ((pos[1::49] - 1) // 49 + 1).tolist()         # rank of each slot's frame
        for ns in slots:
            real = int(mask[start:start + ns].sum())
            faces = [(ranks[start + i] * 10, int(se[start + i]) * 5) for i in range(real)]   # ratio 5*b -> bucket b
            ids.append((ns, faces))
            start += ns
        se2, mask2, idm2, pos2 = clip_meta(ids, 16, source="predict")
        assert (se2 == se).all() and (mask2 == mask).all() and (idm2 == idm).all()
        # positions agree on every slot that holds a face (the generator pads with the identity's own last frame, the
        # executed reference with the clip-wide maximum so far: tests/golden/clip_meta_ref.json)
        real_tok = np.concatenate(([True], np.repeat(mask, 49)))
        assert (pos2[real_tok] == pos[real_tok]).all()
