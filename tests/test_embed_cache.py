"""Host-side tests of the prompt-embedding cache and of encode_prompt's hand-over logic."""
import pytest
import torch


def test_cache_round_trip_and_batching(tmp_path):
    from tpdm_b200.embed_cache import PromptEmbeddingCache

    cache = PromptEmbeddingCache(str(tmp_path / "emb"))
    g = torch.Generator().manual_seed(0)
    prompts = ["a cat", "a dog on a skateboard", ""]
    pe = torch.randn(3, 333, 64, generator=g).to(torch.bfloat16)
    pp = torch.randn(3, 32, generator=g).to(torch.bfloat16)
    cache.put_many(prompts, pe, pp)
    assert all(p in cache for p in prompts) and "a bird" not in cache
    assert sorted(cache.prompts()) == sorted(prompts)
    a, b = cache.get("a dog on a skateboard")
    assert torch.equal(a, pe[1]) and torch.equal(b, pp[1]) and a.dtype == torch.bfloat16      # bit-exact, dtype kept
    A, B = cache.get_batch(["", "a cat", "a cat"], dtype=torch.float32)
    assert A.shape == (3, 333, 64) and torch.equal(A[1], pe[0].float()) and torch.equal(B[0], pp[2].float())
    with pytest.raises(KeyError):
        cache.get("a bird")
    with pytest.raises(ValueError):
        cache.put("bad", torch.zeros(4), torch.zeros(4))
    cache.put("short", torch.zeros(10, 64), torch.zeros(32))
    with pytest.raises(ValueError):
        cache.get_batch(["a cat", "short"])          # ragged token counts are refused


def test_encode_prompt_resolves_through_cache(tmp_path):
    """reference behaviour kept: negative prompt defaults to "", batch-size mismatch raises ValueError
    (modeling_sd3_pnt.py:367-372), num_images_per_prompt repeats each prompt's rows."""
    from tpdm_b200.embed_cache import PromptEmbeddingCache
    from tpdm_b200.modeling_sd3_pnt import SD3PredictNextTimeStepModel

    model = SD3PredictNextTimeStepModel.__new__(SD3PredictNextTimeStepModel)
    torch.nn.Module.__init__(model)
    model.register_buffer("_anchor", torch.zeros(1))
    with pytest.raises(NotImplementedError):
        model.encode_prompt(prompt="a cat")
    cache = PromptEmbeddingCache(str(tmp_path / "emb"))
    for i, p in enumerate(["a cat", "a dog", "", "blurry"]):
        cache.put(p, torch.full((5, 8), float(i)), torch.full((4,), float(i)))
    model.embedding_cache = cache
    pe, ne, pp, npl = model.encode_prompt(prompt=["a cat", "a dog"], num_images_per_prompt=2, device="cpu")
    assert pe.shape == (4, 5, 8) and pe[:, 0, 0].tolist() == [0.0, 0.0, 1.0, 1.0]
    assert ne[:, 0, 0].tolist() == [2.0] * 4 and npl.shape == (4, 4) and pp[2, 0] == 1.0
    pe, ne, _, _ = model.encode_prompt(prompt="a dog", negative_prompt="blurry", device="cpu")
    assert pe.shape == (1, 5, 8) and ne[0, 0, 0] == 3.0
    with pytest.raises(ValueError):
        model.encode_prompt(prompt=["a cat", "a dog"], negative_prompt=["blurry"], device="cpu")
    with pytest.raises(KeyError):
        model.encode_prompt(prompt="unknown", device="cpu")
