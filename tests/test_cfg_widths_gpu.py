"""Parity at the BENCHMARKED widths (VERDICT r1 "Missing 1"): the BASELINE.json configs at full width and reduced depth,
every stage against the fp32 CPU oracle on shared seeded weights.

* cfg2: Whisper-large-v3 encoder widths (d1280 / 20 heads / ffn5120 / 128 mel) + Llama-3.1-8B widths (d4096, 32q / 8kv heads of
  128, ffn14336, V128256), 30 s clip, B=1, S=201 - the exact GEMM / attention shapes `bench.py` times - with 2 + 2 layers.
* cfg3: the same widths, B=2 clips, adapter gradients against torch.autograd on the oracle.
* cfg4: Llama-3.3-70B widths (d8192, 64q / 8kv, ffn28672), 1 layer: prefill + greedy decode against the oracle stepped token
  by token (the oracle re-runs the full sequence each step - no cache on its side).
* cfg5: one ragged cell of the encoder sweep (5 s + 12 s clips, batch-longest padding).

Tolerances: SURVEY.md section 7 (bf16 storage between kernels): encoder <= 1.5e-2, projector <= 2e-2, logits <= 3e-2 relative
(Frobenius) to fp32 math on the same bf16-rounded weights / mel; integer / splice paths bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def wave(i, n):
    return np.random.default_rng(1000 + i).standard_normal(n).astype(np.float32)


def build_width(name, enc_layers, llm_layers, vocab=None, **kw):
    from oracle import model as om
    from ultravox_b200.config import PRESETS, preset
    from ultravox_b200.model import UltravoxModel
    base = PRESETS[name]
    ac = dict(base["audio_config"], encoder_layers=enc_layers)
    tc = dict(base["text_config"], num_hidden_layers=llm_layers)
    if vocab is not None:
        tc["vocab_size"] = vocab
    cfg = preset(name, audio_config=ac, text_config=tc, **kw)
    model = UltravoxModel(cfg, device="cuda").init_random_(seed=42)
    return cfg, model, om.state_dict_fp32(model), om.shapes_from_config(cfg)


def prompt(cfg, lens_samples, seed=7):
    """SURVEY 8d layout: 8 ids, placeholders, 5 ids; one sequence per clip."""
    from oracle import logmel as ol
    waves = [wave(i, n) for i, n in enumerate(lens_samples)]
    padded, frames = ol.pad_batch(waves)
    g = torch.Generator().manual_seed(seed)
    tok = [int(-(-int(f) // 16)) for f in frames]
    S = 8 + max(tok) + 5
    ids = torch.randint(0, min(cfg.vocab_size, 128000), (len(waves), S), generator=g)
    batch = dict(input_ids=ids, audio_token_start_idx=torch.tensor([8] * len(waves)),
                 audio_lens=torch.tensor([int(f) for f in frames]), audio_token_len=torch.tensor(tok, dtype=torch.int32),
                 audio_batch_size=torch.ones(len(waves), dtype=torch.int64))
    return padded, batch


def test_cfg2_widths_prefill_every_stage_vs_oracle():
    from oracle import logmel as ol, model as om
    from ultravox_b200 import ops
    from ultravox_b200.engine import PrefillEngine
    cfg, model, sd, sh = build_width("v0_5_8b", 2, 2)
    padded, batch = prompt(cfg, [16000 * 30])
    assert batch["input_ids"].shape == (1, 201) and batch["audio_token_len"].tolist() == [188]
    mel_ref = torch.from_numpy(ol.log_mel(padded, sh.n_mels))
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    assert tuple(mel.shape) == (1, 128, 3000) and float((mel.cpu() - mel_ref).abs().max()) < 2e-3
    mel_b = mel.cpu().to(BF).float()                      # what the GPU path consumes
    st = {}
    ref_logits, _ = om.forward(sd, sh, batch["input_ids"], mel_b, batch["audio_token_start_idx"], batch["audio_lens"],
                               batch["audio_token_len"], batch["audio_batch_size"], last_only=True, stages=st)
    enc = model.encode_audio(ops.mel_to_timemajor(mel), batch["audio_lens"])
    assert tuple(enc.shape) == (1, 1500, 1280)
    r_enc = rel(enc, st["encoder"])
    aud = model.project_audio(enc)
    assert tuple(aud.shape) == (1, 188, 4096)
    r_proj = rel(aud, st["projector"])
    emb = model._prepare_audio_embeds(batch["input_ids"].cuda(), mel, batch["audio_token_start_idx"], batch["audio_lens"],
                                      batch["audio_token_len"], batch["audio_batch_size"])
    table = model.language_model.model.embed_tokens.weight
    ids = batch["input_ids"][0].cuda()
    assert torch.equal(emb[0, 8:196], aud[0, :188])                              # splice rows: bit-exact copies
    assert torch.equal(emb[0, :8], table[ids[:8]]) and torch.equal(emb[0, 196:], table[ids[196:]])
    r_emb = rel(emb, st["inputs_embeds"])
    out = model(audio_values=mel, logits_to_keep=1, **{k: v.cuda() for k, v in batch.items()})
    got = out.logits.view(1, 1, -1).cpu()
    r_log = rel(got, ref_logits)
    top = got.view(-1).topk(5).indices.tolist()
    print(f"cfg2 widths: encoder {r_enc:.3e} projector {r_proj:.3e} embeds {r_emb:.3e} last-row logits {r_log:.3e}")
    assert r_enc < 1.5e-2 and r_proj < 2e-2 and r_emb < 2e-2 and r_log < 3e-2
    assert int(ref_logits.view(-1).argmax()) in top
    # the graph-captured engine bench.py times produces the same token / logits as the eager forward
    eng = PrefillEngine(model, 16000 * 30, batch["input_ids"], batch["audio_token_start_idx"], batch["audio_token_len"],
                        batch["audio_batch_size"])
    tok = eng.run_e2e(torch.from_numpy(padded).pin_memory()).clone()
    assert int(tok[0]) == int(got.view(-1).argmax())
    assert rel(eng.logits, ref_logits.view(1, -1)) < 3e-2


def test_cfg2_widths_all_rows_logits_and_loss():
    """All-row logits ([201, 128256] fp32 through the tensor-core head) + CE loss at width, 1 + 1 layers."""
    from oracle import model as om
    from ultravox_b200 import ops
    cfg, model, sd, sh = build_width("v0_5_8b", 1, 1)
    padded, batch = prompt(cfg, [16000 * 30])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    labels = batch["input_ids"].clone()
    labels[:, :-5] = -100
    out = model(audio_values=mel, labels=labels.cuda(), **{k: v.cuda() for k, v in batch.items()})
    ref_logits, ref_loss = om.forward(sd, sh, batch["input_ids"], mel.cpu().to(BF).float(), batch["audio_token_start_idx"],
                                      batch["audio_lens"], batch["audio_token_len"], batch["audio_batch_size"], labels=labels)
    r = rel(out.logits, ref_logits)
    assert r < 3e-2, r
    assert abs(float(out.loss) - float(ref_loss)) < 3e-2 * max(1.0, abs(float(ref_loss)))


def test_cfg3_widths_adapter_gradients_vs_oracle_autograd():
    from oracle import model as om
    from ultravox_b200 import ops
    from ultravox_b200.training import AdapterTrainer
    cfg, model, sd, sh = build_width("v0_5_8b", 1, 2, vocab=32000)
    padded, batch = prompt(cfg, [16000 * 30, 16000 * 30])
    labels = batch["input_ids"].clone()
    labels[:, :-5] = -100
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    tr = AdapterTrainer(model, lr=1e-3)
    loss = tr.forward_backward(audio_values=mel, labels=labels, **batch)
    names = ["multi_modal_projector." + n + ".weight" for n in ("ln_pre", "linear_1", "ln_mid", "linear_2")]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    _, ref_loss = om.forward(sd, sh, batch["input_ids"], mel.cpu().to(BF).float(), batch["audio_token_start_idx"],
                             batch["audio_lens"], batch["audio_token_len"], batch["audio_batch_size"], labels=labels)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) < 3e-2 * max(1.0, abs(float(ref_loss)))
    for n in names:
        got = tr.grad_view(n.split(".")[1])
        r = rel(got, sd[n].grad)
        cos = float(F.cosine_similarity(got.float().cpu().flatten(), sd[n].grad.flatten(), dim=0))
        print(f"cfg3 widths: {n} rel {r:.3e} cos {cos:.5f}")
        assert r < 8e-2 and cos > 0.995, (n, r, cos)


def test_cfg4_widths_prefill_and_decode_vs_stepwise_oracle():
    """70B widths, 1 layer: the CUDA path (KV cache + per-token steps) against the oracle re-forwarding the whole sequence for
    every new token.  Random-init logits are nearly tied, so the oracle is teacher-forced with the CUDA tokens and each CUDA
    token must sit in the oracle's top-5 with a small logit gap (and the last-row logits agree within tolerance)."""
    from oracle import model as om
    from ultravox_b200 import ops
    from ultravox_b200.engine import DecodeEngine
    cfg, model, sd, sh = build_width("v0_5_70b", 1, 1, vocab=32000)
    padded, batch = prompt(cfg, [16000 * 30])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    kw = {k: v.cuda() for k, v in batch.items()}
    n_new = 4
    seq = model.generate(audio_values=mel, max_new_tokens=n_new, **kw)
    S = batch["input_ids"].shape[1]
    assert seq.shape == (1, S + n_new)
    emb_ref = {}
    om.forward(sd, sh, batch["input_ids"], mel.cpu().to(BF).float(), batch["audio_token_start_idx"], batch["audio_lens"],
               batch["audio_token_len"], batch["audio_batch_size"], last_only=True, stages=emb_ref)
    cur = emb_ref["inputs_embeds"]
    table = sd["language_model.model.embed_tokens.weight"]
    for t in range(n_new):
        ref = om.llama_forward(sd, sh, cur, last_only=True).view(-1)
        tok = int(seq[0, S + t])
        assert tok in ref.topk(5).indices.tolist(), t
        assert float(ref.max() - ref[tok]) < 3e-2 * float(ref.abs().max()), t
        cur = torch.cat([cur, table[tok][None, None]], dim=1)
    # the graph-captured decode engine emits the same tokens as generate()
    emb = model._prepare_audio_embeds(kw["input_ids"], mel, batch["audio_token_start_idx"], batch["audio_lens"],
                                      batch["audio_token_len"], batch["audio_batch_size"])
    de = DecodeEngine(model, 1, S + n_new + 2)
    toks = [int(de.prefill(emb)[0])]
    for _ in range(n_new - 1):
        toks.append(int(de.step().view(-1)[0]))
    assert toks == seq[0, S:].tolist()


def test_cfg5_widths_ragged_encoder_cell():
    from oracle import logmel as ol, model as om
    from ultravox_b200 import ops
    cfg, model, sd, sh = build_width("v0_5_8b", 2, 1, vocab=32000)
    padded, batch = prompt(cfg, [16000 * 5, 16000 * 12])
    mel = ops.logmel(torch.from_numpy(padded).cuda(), sh.n_mels)
    assert tuple(mel.shape) == (2, 128, 1200)
    enc_ref = om.whisper_encoder(sd, sh, mel.cpu().to(BF).float(), batch["audio_lens"])
    enc = model.encode_audio(ops.mel_to_timemajor(mel), batch["audio_lens"])
    for i, n in enumerate(batch["audio_lens"].tolist()):
        valid = (n - 1) // 2 + 1
        r = rel(enc[i, :valid], enc_ref[i, :valid])
        assert r < 1.5e-2, (i, r)
