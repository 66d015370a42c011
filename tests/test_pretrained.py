"""Pretrained-weight handling of the BERT stacks (reference: from_pretrained('bert-base-uncased') in training/networks_detr.py:92,124):
a local checkpoint is loaded by parameter name; a missing one is an error unless synthetic weights were requested."""
import os

import pytest
import torch


def _tiny_cfg():
    from layoutdetr_b200.training import med
    cfg = med.BertConfig.default()
    cfg.num_hidden_layers = 1
    cfg.hidden_size = 128
    cfg.intermediate_size = 256
    cfg.num_attention_heads = 2
    cfg.encoder_width = 128
    cfg.vocab_size = 64
    cfg.max_position_embeddings = 16
    return cfg


def test_missing_checkpoint_is_an_error_without_opt_in(monkeypatch, tmp_path):
    from layoutdetr_b200.training import med
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_WEIGHTS", "0")
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_TOKENIZER", "0")
    with pytest.raises(FileNotFoundError):
        med.BertModel.from_pretrained(str(tmp_path / "nowhere"), config=_tiny_cfg(), add_pooling_layer=False)
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_WEIGHTS", "1")
    med.BertModel.from_pretrained(str(tmp_path / "nowhere"), config=_tiny_cfg(), add_pooling_layer=False)      # explicit opt-in: random init


def test_local_checkpoint_is_loaded_by_name(monkeypatch, tmp_path):
    from layoutdetr_b200.training import med
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_WEIGHTS", "0")
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_TOKENIZER", "0")
    src = med.BertLMHeadModel(_tiny_cfg())
    sd = {}
    for k, v in src.state_dict().items():                     # the hub checkpoint's naming: 'bert.' prefix, TF-style LayerNorm names
        if "crossattention" in k or "position_ids" in k:
            continue
        k2 = k.replace("LayerNorm.weight", "LayerNorm.gamma").replace("LayerNorm.bias", "LayerNorm.beta")
        sd[k2] = v.clone()
    d = tmp_path / "ckpt"
    d.mkdir()
    torch.save(sd, d / "pytorch_model.bin")
    enc = med.BertModel.from_pretrained(str(d), config=_tiny_cfg(), add_pooling_layer=False)
    for k, v in enc.state_dict().items():
        if "crossattention" in k or "position_ids" in k:
            continue
        assert torch.equal(v, src.state_dict()["bert." + k]), k
    dec = med.BertLMHeadModel.from_pretrained(str(d), config=_tiny_cfg())
    assert torch.equal(dec.cls.predictions.transform.dense.weight, src.cls.predictions.transform.dense.weight)
    assert dec.cls.predictions.decoder.weight is dec.bert.embeddings.word_embeddings.weight


def test_tokenizer_needs_vocabulary_or_opt_in(monkeypatch):
    from layoutdetr_b200.training import networks_detr as nd
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_TOKENIZER", "1")
    assert len(nd.init_tokenizer()) == 30524
    monkeypatch.setenv("LAYOUTDETR_SYNTHETIC_TOKENIZER", "0")
    monkeypatch.setenv("HF_HUB_OFFLINE", "1")
    try:
        tok = nd.init_tokenizer()            # a machine with the vocabulary cached gets the real tokenizer
        assert len(tok) == 30524
    except RuntimeError as e:
        assert "LAYOUTDETR_SYNTHETIC_TOKENIZER" in str(e)
