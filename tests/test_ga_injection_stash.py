"""Gradient accumulation wrapper, backward injection hooks, memory stashing (CPU)."""
import pytest
import torch
from torch import nn


class _Pipe:
    def __init__(self, model, opt):
        self._model, self._optimizer = model, opt

    def progress(self, it):
        x, y = next(it)
        self._optimizer.zero_grad()
        loss = ((self._model(x) - y) ** 2).mean()
        loss.backward()
        self._optimizer.step()
        return loss


def test_gradient_accumulation_matches_big_batch():
    from torchrec_b200.parallel.train_pipeline.gradient_accumulation import GradientAccumulationConfig, GradientAccumulationWrapper

    torch.manual_seed(0)
    m1, m2 = nn.Linear(4, 1), nn.Linear(4, 1)
    m2.load_state_dict(m1.state_dict())
    o1, o2 = torch.optim.SGD(m1.parameters(), lr=0.1), torch.optim.SGD(m2.parameters(), lr=0.1)
    data = [(torch.randn(8, 4), torch.randn(8, 1)) for _ in range(9)]
    ga = GradientAccumulationWrapper(_Pipe(m1, o1), o1, m1, GradientAccumulationConfig(num_steps=4, num_warmup_steps=1))
    it = iter(data)
    for _ in range(9):
        ga.progress(it)
    # golden: warm-up step alone, then two windows of 4 micro-batches (summed gradients)
    o2.zero_grad(); ((m2(data[0][0]) - data[0][1]) ** 2).mean().backward(); o2.step()
    for w in range(2):
        o2.zero_grad()
        for x, y in data[1 + 4 * w : 5 + 4 * w]:
            ((m2(x) - y) ** 2).mean().backward()
        o2.step()
    torch.testing.assert_close(m1.weight, m2.weight)
    assert ga.current_step == 9
    with pytest.raises(ValueError):
        GradientAccumulationConfig(num_steps=0)


def test_backward_injection_sites():
    from torchrec_b200.parallel.train_pipeline.backward_injection import InjectionSite, InjectionTargetType, register_backward_hook

    model = nn.Sequential(nn.Linear(4, 4), nn.ReLU(), nn.Linear(4, 2))
    fired = []
    h1 = register_backward_hook(InjectionSite(fqn="0", target_type=InjectionTargetType.ACTIVATION), model, lambda g: fired.append(("act", tuple(g.shape))))
    h2 = register_backward_hook(InjectionSite(fqn="2", target_type=InjectionTargetType.PARAM_GRAD, hook_position=0.0), model, lambda g: fired.append(("param", tuple(g.shape))))
    model(torch.randn(3, 4)).sum().backward()
    assert ("act", (3, 4)) in fired and ("param", (2, 4)) in fired
    assert fired.index(("param", (2, 4))) < fired.index(("act", (3, 4)))  # later layers finish first in backward
    h1.remove(); h2.remove()
    fired.clear()
    model(torch.randn(3, 4)).sum().backward()
    assert fired == []
    with pytest.raises(ValueError):
        register_backward_hook(InjectionSite(fqn="nope"), model, lambda g: None)


def test_memory_stashing_round_trip():
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection
    from torchrec_b200.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec_b200.optim.rowwise_adagrad import RowWiseAdagrad
    from torchrec_b200.parallel import sharding_plan as sp
    from torchrec_b200.parallel.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec_b200.parallel.memory_stashing import MemoryStashingManager as M
    from torchrec_b200.parallel.model_parallel import DistributedModelParallel
    from torchrec_b200.parallel.types import ShardingPlan
    from torchrec_b200.sparse import KeyedJaggedTensor

    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t", embedding_dim=8, num_embeddings=64, feature_names=["f"])])
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": 0.1})
    plan = sp.construct_module_sharding_plan(ebc, {"t": sp.table_wise(rank=0)}, sharder=EmbeddingBagCollectionSharder(), world_size=1, local_size=1, device_type="cpu")

    class W(nn.Module):
        def __init__(self):
            super().__init__()
            self.ebc = ebc

        def forward(self, k):
            return self.ebc(k).values()

    model = DistributedModelParallel(W(), device=torch.device("cpu"), plan=ShardingPlan({"ebc": plan}), sharders=[EmbeddingBagCollectionSharder()])
    kjt = KeyedJaggedTensor(keys=["f"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([2, 1]))
    model(kjt).sum().backward()
    before = model(kjt).detach().clone()
    tbe = model.module.ebc.engine._tbes[0]
    st_before = tbe.state1.clone()
    M.reset()  # the manager is process-wide: leftovers of other tests (benchmark pipelines with stashing) must not count here
    n = M.stash_embedding_weights(model) + M.stash_optimizer_state(model)
    assert n > 0 and M.stashed_bytes() == n and tbe.weights.untyped_storage().nbytes() == 0
    M.restore_embedding_weights(); M.restore_optimizer_state()
    assert M.stashed_bytes() == 0
    torch.testing.assert_close(model(kjt).detach(), before)
    torch.testing.assert_close(tbe.state1, st_before)
    M.set_delay_stash(True)
    M.stash_optimizer_state(model)
    assert M.stashed_bytes() == 0
    M.execute_pending_stashes()
    assert M.stashed_bytes() > 0
    M.reset()
    torch.testing.assert_close(tbe.state1, st_before)
