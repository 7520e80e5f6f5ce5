"""Pins the CPU oracle (oracle/) against (a) fixtures produced by the reference's own code run in the build
container (tests/golden, scripts/make_golden.py), (b) the literals of the reference's own tests, (c) the third-party
transformers modules that hold the path's arithmetic.  No GPU needed."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import logmel, model as om, processing as op

G = os.path.join(os.path.dirname(__file__), "golden")


def stub_tokenize(parts):
    return [[(sum(map(ord, w)) * 31 + len(w)) % 100000 for w in p.split()] for p in parts]


def wave(i, n):
    return np.random.default_rng(1000 + i).standard_normal(n).astype(np.float32)


@pytest.fixture(scope="module")
def cases():
    return json.load(open(os.path.join(G, "processor_cases.json")))


# ---------------------------------------------------------------- integer path: fixtures made by the reference class
def test_processor_matches_reference_fixtures(cases):
    for c in cases["cases"]:
        audios = [wave(i, n) for i, n in enumerate(c["sample_counts"])]
        got = op.process(c["text"], audios, stub_tokenize, 128009, n_mels=80, include_audio_num_chunks=True,
                         with_mel=False)
        exp = c["out"]
        for key in ("input_ids", "attention_mask", "audio_lens", "audio_token_len", "audio_token_start_idx",
                    "audio_batch_size", "audio_num_chunks"):
            if key in exp:
                assert np.asarray(got[key]).tolist() == exp[key], (c["name"], key)
            else:
                assert key not in got, (c["name"], key)


def test_processor_audio_values_shape_and_chunks(cases):
    by = {c["name"]: c for c in cases["cases"]}
    for name in ("one_1s", "one_35s", "three_overflow", "short_0", "short_321"):
        c = by[name]
        audios = [wave(i, n) for i, n in enumerate(c["sample_counts"])]
        got = op.process(c["text"], audios, stub_tokenize, 128009, n_mels=80)
        assert list(got["audio_values"].shape) == c["out"]["audio_values_shape"], name


def test_processor_errors_match_reference(cases):
    for e in cases["errors"]:
        audios = [wave(i, n) for i, n in enumerate(e["sample_counts"])]
        if e["raises"]:
            with pytest.raises(ValueError) as ei:
                op.process(e["text"], audios, stub_tokenize, 128009, with_mel=False)
            assert str(ei.value) == e["msg"]
        else:
            op.process(e["text"], audios, stub_tokenize, 128009, with_mel=False)


def test_reference_test_literals():
    """Tokenizer-independent numbers asserted by ref ultravox_processing_test.py:46-137 and infer_test.py:72-109."""
    sr = 16000
    r = op.process("a b c <|audio|>", [wave(0, sr)], stub_tokenize, 1, with_mel=False)
    assert r["audio_lens"].tolist() == [100] and r["audio_token_len"].tolist() == [7]
    assert r["audio_token_start_idx"].tolist() == [3] and r["input_ids"].shape[1] == 3 + 7
    r = op.process("a b c <|audio|>", [wave(0, 35 * sr)], stub_tokenize, 1, with_mel=False)
    assert r["audio_lens"].tolist() == [3000, 500] and r["audio_token_len"].tolist() == [188, 32]
    assert r["audio_token_start_idx"].tolist() == [3, 3 + 188] and r["audio_batch_size"].tolist() == [2]
    r = op.process("a b c <|audio|> and d <|audio|> and e <|audio|>", [wave(0, sr), wave(1, 35 * sr), wave(2, 10 * sr)],
                   stub_tokenize, 1, include_audio_num_chunks=True, with_mel=False)
    assert r["audio_lens"].tolist() == [100, 3000, 500, 1000] and r["audio_token_len"].tolist() == [7, 188, 32, 63]
    assert r["audio_token_start_idx"].tolist() == [3, 12, 200, 234] and r["audio_num_chunks"].tolist() == [1, 2, 1]
    # infer_test.py:72-86: 60 s -> 2 chunks of 188 tokens starting at 8 and 196, 389 ids
    r = op.process("1 2 3 4 5 6 7 8 <|audio|> 9 10 11 12 13", [wave(0, 60 * sr)], stub_tokenize, 1, with_mel=False)
    assert r["audio_token_len"].tolist() == [188, 188] and r["audio_token_start_idx"].tolist() == [8, 196]
    assert r["input_ids"].shape[1] == 389
    for n in (0, 1, 159, 160, 161, 319, 320, 321):   # ref :177-186
        r = op.process("<|audio|>", [wave(0, n)], stub_tokenize, 1, n_mels=80)
        assert r["audio_lens"][0] == r["audio_values"][0].shape[-1]


def test_collator_matches_reference_fixtures(cases):
    sr = 16000
    for c in cases["collator"]:
        feats = []
        for text, n, i in (("Test with <|audio|>", sr, 0), ("Other longer text with <|audio|> more", 35 * sr, 1)):
            f = op.process(text, [wave(i, n)], stub_tokenize, 128009, n_mels=80)
            f["input_ids"], f["attention_mask"] = f["input_ids"][0], f["attention_mask"][0]
            feats.append(f)
        got = op.collate(feats, pad_id=128009, padding_side=c["padding_side"])
        exp = c["out"]
        for key in ("input_ids", "attention_mask", "audio_lens", "audio_token_len", "audio_token_start_idx",
                    "audio_batch_size"):
            assert np.asarray(got[key]).tolist() == exp[key], (c["padding_side"], key)
        assert list(got["audio_values"].shape) == exp["audio_values_shape"]
        assert abs(float(got["audio_values"].astype(np.float64).sum()) - exp["audio_values_sum"]) < 0.5


# ---------------------------------------------------------------- log-mel: HF feature extractor fixtures + live
def test_logmel_matches_hf_fixture():
    z = np.load(os.path.join(G, "logmel_hf.npz"))
    sr = 16000
    t = np.arange(2 * sr) / sr
    tone = (0.1 * np.sin(2 * np.pi * 440 * t)).astype(np.float32)
    tone[int(0.75 * len(tone)):] = 0
    waves = [wave(0, sr), tone, wave(2, sr // 2 + 33)]
    padded, frame_lens = logmel.pad_batch(waves)
    for n_mels, tag in ((80, "m80"), (128, "m128")):
        got = logmel.log_mel(padded, n_mels)
        ref = z[f"{tag}_mel"]
        assert got.shape == ref.shape
        # fp32 pocketfft (reference) vs float64 (oracle): near-floor bins of the tonal clip carry the fp32 noise
        assert np.abs(got - ref).max() < 2e-3 and np.sqrt(((got - ref) ** 2).mean()) < 1e-4
        assert z[f"{tag}_mask"].sum(-1).tolist() == frame_lens.tolist()
        got32 = logmel.log_mel(padded, n_mels, dtype=np.float32)
        assert np.abs(got32 - ref).max() < 2e-3


def test_mel_filters_match_hf():
    transformers = pytest.importorskip("transformers")
    for n in (80, 128):
        hf = transformers.WhisperFeatureExtractor(feature_size=n).mel_filters
        assert np.array_equal(logmel.mel_filter_bank(n).astype(np.float32), hf.astype(np.float32))


# ---------------------------------------------------------------- projector / stack / splice: reference classes
@pytest.mark.parametrize("tag,ln_mid", [("v05", True), ("v04", False)])
def test_projector_matches_reference_fixture(tag, ln_mid):
    z = np.load(os.path.join(G, "projector_ref.npz"))
    sd = {"multi_modal_projector." + k[len(tag) + 3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(f"{tag}_w_")}
    sh = om.Shapes(enc_d=64, stack=8, proj_hidden=128, proj_ln_mid=ln_mid, d=96)
    x = torch.from_numpy(z[f"{tag}_in"])
    assert torch.equal(om.stack_frames(x, 8), torch.from_numpy(z[f"{tag}_stacked"]))
    got = om.projector(sd, sh, x)
    assert torch.allclose(got, torch.from_numpy(z[f"{tag}_out"]), rtol=1e-5, atol=1e-6)


def test_splice_matches_reference_fixture():
    z = np.load(os.path.join(G, "projector_ref.npz"))
    got = om.splice(torch.from_numpy(z["splice_emb"]).clone(), torch.from_numpy(z["splice_audio"]),
                    torch.from_numpy(z["splice_start"]), torch.from_numpy(z["splice_len"]), torch.from_numpy(z["splice_abs"]))
    assert torch.equal(got, torch.from_numpy(z["splice_out"]))


def test_latency_mask_matches_reference_fixture():
    z = np.load(os.path.join(G, "latency_mask.npz"))
    m = om.encoder_masks(torch.tensor([3000]), 1500, torch.float32, int(z["block"]), 3000)
    full = om.encoder_masks(torch.tensor([3000]), 3000, torch.float32, int(z["block"]), 3000) if False else None
    del full
    allowed = (m[0, 0, ::50, ::50] == 0).numpy()
    assert np.array_equal(allowed, z["allowed"][:30, :30])
    # values/dtype semantic of ref ultravox_model_test.py:25-68: 0 where allowed, finfo.min elsewhere
    assert m.dtype == torch.float32 and float(m.min()) == torch.finfo(torch.float32).min and float(m.max()) == 0.0


# ---------------------------------------------------------------- third-party arithmetic: transformers modules
def _rand_sd(shapes, seed=0, std=0.05):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.randn(*s, generator=g) * std for k, s in shapes.items()}


def test_whisper_encoder_matches_transformers():
    transformers = pytest.importorskip("transformers")
    from transformers.models.whisper import modeling_whisper as mw
    cfg = transformers.WhisperConfig(d_model=64, encoder_layers=2, encoder_attention_heads=2, encoder_ffn_dim=128,
                                     num_mel_bins=80, max_source_positions=1500, decoder_layers=1,
                                     decoder_attention_heads=2, decoder_ffn_dim=64)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    enc = mw.WhisperEncoder(cfg).eval().float()
    sd = {"audio_tower." + k: v for k, v in enc.state_dict().items()}
    sh = om.Shapes(n_mels=80, enc_d=64, enc_layers=2, enc_heads=2, enc_ffn=128)
    mel = torch.randn(2, 80, 3000)
    with torch.no_grad():
        ref = enc(mel).last_hidden_state
        got = om.whisper_encoder(sd, sh, mel, None)
    assert torch.allclose(got, ref, rtol=1e-4, atol=1e-5)
    # ragged (audio_len mask) against the layer modules driven exactly like the reference wrapper drives them
    mel2 = torch.randn(2, 80, 200)
    alen = torch.tensor([200, 61])
    with torch.no_grad():
        h = torch.nn.functional.gelu(enc.conv1(mel2))
        h = torch.nn.functional.gelu(enc.conv2(h)).permute(0, 2, 1)
        h = h + enc.embed_positions.weight[: h.size(-2)]
        m = om.encoder_masks(alen, h.shape[1], h.dtype, None)
        for layer in enc.layers:
            out = layer(h, m)
            h = out[0] if isinstance(out, tuple) else out
        ref2 = enc.layer_norm(h)
        got2 = om.whisper_encoder(sd, sh, mel2, alen)
    assert torch.allclose(got2, ref2, rtol=1e-4, atol=1e-5)


def test_llama_matches_transformers():
    transformers = pytest.importorskip("transformers")
    cfg = transformers.LlamaConfig(hidden_size=128, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
                                   intermediate_size=256, vocab_size=512, head_dim=32, rms_norm_eps=1e-5,
                                   max_position_embeddings=131072, rope_theta=500000.0,
                                   rope_scaling=dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0,
                                                     high_freq_factor=4.0, original_max_position_embeddings=8192),
                                   tie_word_embeddings=False)
    cfg._attn_implementation = "eager"
    torch.manual_seed(0)
    lm = transformers.LlamaForCausalLM(cfg).eval().float()
    sd = {"language_model." + k: v for k, v in lm.state_dict().items()}
    sh = om.Shapes(d=128, layers=2, heads=4, kv_heads=2, head_dim=32, ffn=256, vocab=512)
    emb = torch.randn(2, 23, 128)
    labels = torch.randint(0, 512, (2, 23))
    labels[:, :15] = -100
    with torch.no_grad():
        ref = lm(inputs_embeds=emb, labels=labels)
        got = om.llama_forward(sd, sh, emb)
        loss = om.causal_lm_loss(got, labels)
    assert torch.allclose(got, ref.logits, rtol=1e-4, atol=1e-4)
    assert abs(float(loss) - float(ref.loss)) < 1e-4
    am = torch.ones(2, 23, dtype=torch.long)
    am[1, 18:] = 0
    with torch.no_grad():
        ref2 = lm(inputs_embeds=emb, attention_mask=am).logits
        got2 = om.llama_forward(sd, sh, emb, attention_mask=am)
    assert torch.allclose(got2[0], ref2[0], rtol=1e-4, atol=1e-4)
    assert torch.allclose(got2[1, :18], ref2[1, :18], rtol=1e-4, atol=1e-4)
