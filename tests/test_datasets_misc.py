"""Datasets (Criteo TSV/binary, MovieLens), DeepFM model, KJT validator, TensorDict interop, packed tensor types."""
import numpy as np
import os

import pytest
import torch

from torchrec_b200.sparse import JaggedTensor, KeyedJaggedTensor


def _write_days(d, n_rows=50, days=2):
    from torchrec_b200.datasets.criteo import BinaryCriteoUtils

    rng = np.random.default_rng(0)
    for day in range(days):
        rows = []
        for i in range(n_rows):
            ints = [str(rng.integers(0, 100)) if rng.random() > 0.1 else "" for _ in range(13)]
            cats = [format(rng.integers(0, 2**32), "x") if rng.random() > 0.1 else "" for _ in range(26)]
            rows.append("\t".join([str(i % 2)] + ints + cats))
        (d / f"day_{day}").write_text("\n".join(rows) + "\n")
        BinaryCriteoUtils.tsv_to_npys(str(d / f"day_{day}"), str(d / f"day_{day}_dense.npy"), str(d / f"day_{day}_sparse.npy"), str(d / f"day_{day}_labels.npy"))


def test_criteo_tsv_and_binary(tmp_path):
    from torchrec_b200.datasets.criteo import DEFAULT_CAT_NAMES, BinaryCriteoUtils, InMemoryBinaryCriteoIterDataPipe, criteo_terabyte

    _write_days(tmp_path)
    rows = list(criteo_terabyte([str(tmp_path / "day_0")]))
    assert len(rows) == 50 and set(rows[0]) >= {"label", "int_0", "cat_25"}
    assert BinaryCriteoUtils.get_shape_from_npy(str(tmp_path / "day_0_sparse.npy")) == (50, 26)
    ranges, rem = BinaryCriteoUtils.get_file_row_ranges_and_remainder([50, 50], rank=1, world_size=3)
    assert ranges == {0: (34, 49), 1: (0, 16)} and rem == 1
    paths = lambda kind: [str(tmp_path / f"day_{k}_{kind}.npy") for k in range(2)]
    seen = 0
    for r in range(2):
        dp = InMemoryBinaryCriteoIterDataPipe("train", paths("dense"), paths("sparse"), paths("labels"), batch_size=16, rank=r, world_size=2, hashes=[1000] * 26)
        bs = list(dp)
        assert len(bs) == len(dp) == 4 and bs[0].sparse_features.keys() == DEFAULT_CAT_NAMES
        assert bs[0].dense_features.shape == (16, 13) and int(bs[0].sparse_features.values().max()) < 1000
        assert bs[0].sparse_features.lengths().sum() == 16 * 26
        seen += sum(b.labels.numel() for b in bs)
    assert seen == 100
    val = InMemoryBinaryCriteoIterDataPipe("val", paths("dense"), paths("sparse"), paths("labels"), batch_size=8, rank=0, world_size=1)
    test = InMemoryBinaryCriteoIterDataPipe("test", paths("dense"), paths("sparse"), paths("labels"), batch_size=8, rank=0, world_size=1)
    assert sum(b.labels.numel() for b in val) == 25 and sum(b.labels.numel() for b in test) == 25
    # first row of day_1 == first val row
    torch.testing.assert_close(next(iter(val)).dense_features[0], torch.from_numpy(np.load(paths("dense")[1])[0]))
    BinaryCriteoUtils.sparse_to_contiguous(paths("sparse"), str(tmp_path / "contig"), frequency_threshold=1)
    c = np.load(str(tmp_path / "contig" / "day_0_sparse_contig_freq.npy"))
    assert c.min() >= 0 and c.max() <= 100 + 2


def test_movielens(tmp_path):
    from torchrec_b200.datasets.movielens import movielens_20m

    (tmp_path / "ratings.csv").write_text("userId,movieId,rating,timestamp\n1,10,4.5,100\n2,11,3.0,101\n")
    (tmp_path / "movies.csv").write_text("movieId,title,genres\n10,Foo (1999),Drama\n11,Bar,Comedy|Action\n")
    rows = list(movielens_20m(str(tmp_path), include_movies_data=True))
    assert rows[0]["userId"] == 1 and rows[0]["rating"] == 4.5 and rows[1]["genres"] == "Comedy|Action"


def test_deepfm_model():
    from torchrec_b200.models.deepfm import SimpleDeepFMNN
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection

    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t1", embedding_dim=8, num_embeddings=100, feature_names=["f1", "f3"]),
                                  EmbeddingBagConfig(name="t2", embedding_dim=8, num_embeddings=100, feature_names=["f2"])])
    m = SimpleDeepFMNN(num_dense_features=10, embedding_bag_collection=ebc, hidden_layer_size=20, deep_fm_dimension=5)
    kjt = KeyedJaggedTensor.from_offsets_sync(keys=["f1", "f3", "f2"], values=torch.tensor([1, 2, 4, 5, 4, 3, 2, 9, 1, 2]), offsets=torch.tensor([0, 2, 4, 6, 8, 9, 10]))
    out = m(torch.rand(2, 10), kjt)
    assert out.shape == (2, 1) and bool(((out > 0) & (out < 1)).all())
    out.sum().backward()
    assert ebc.embedding_bags["t1"].weight.grad is not None


def test_kjt_validator():
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.sparse.jagged_tensor_validator import validate_keyed_jagged_tensor

    good = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([1, 0, 2, 0]))
    assert validate_keyed_jagged_tensor(good)
    cfgs = [EmbeddingBagConfig(name="t", embedding_dim=4, num_embeddings=3, feature_names=["a", "b"])]
    assert validate_keyed_jagged_tensor(good, cfgs) is False  # id 3 out of range
    with pytest.raises(ValueError, match="Sum of lengths"):
        validate_keyed_jagged_tensor(KeyedJaggedTensor(keys=["a"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([1, 1])))
    with pytest.raises(ValueError, match="unique"):
        validate_keyed_jagged_tensor(KeyedJaggedTensor(keys=["a", "a"], values=torch.tensor([1]), lengths=torch.tensor([1, 0])))
    with pytest.raises(ValueError, match="weights size"):
        validate_keyed_jagged_tensor(KeyedJaggedTensor(keys=["a"], values=torch.tensor([1, 2]), lengths=torch.tensor([2]), weights=torch.tensor([1.0])))
    with pytest.raises(ValueError, match="first offset"):
        validate_keyed_jagged_tensor(KeyedJaggedTensor(keys=["a"], values=torch.tensor([1, 2]), offsets=torch.tensor([1, 2])))


def test_tensor_dict_interop_and_uintx():
    from torchrec_b200.sparse.tensor_dict import maybe_td_to_kjt
    from torchrec_b200.tensor_types import UInt2Tensor, UInt4Tensor

    td = {"a": JaggedTensor(values=torch.tensor([1, 2, 3]), lengths=torch.tensor([2, 1])), "b": [torch.tensor([7]), torch.tensor([], dtype=torch.int64)]}
    kjt = maybe_td_to_kjt(td)
    assert kjt.keys() == ["a", "b"] and kjt.values().tolist() == [1, 2, 3, 7] and kjt.lengths().tolist() == [2, 1, 1, 0]
    assert maybe_td_to_kjt(kjt) is kjt
    v = torch.randint(0, 16, (3, 8))
    p = UInt4Tensor.pack(v)
    assert p.shape == (3, 8) and p.elem.shape == (3, 4) and torch.equal(p.unpack().long(), v)
    v2 = torch.randint(0, 4, (2, 16))
    p2 = UInt2Tensor.pack(v2)
    assert p2.elem.shape == (2, 4) and torch.equal(p2.unpack().long(), v2) and p2[0].shape == (16,) and isinstance(p2, torch.Tensor)
    # a real tensor subclass: row / whole-byte column slices, clone / detach / copy_ / byte view keep working on the packed storage
    assert torch.equal(p[1:3, 2:6].unpack().long(), v[1:3, 2:6]) and type(p[1:3]) is UInt4Tensor
    q = p.clone().detach()
    q.copy_(UInt4Tensor.pack((v + 1) % 16))
    assert torch.equal(q.unpack().long(), (v + 1) % 16) and torch.equal(p.unpack().long(), v) and p.view(torch.uint8).shape == (3, 4)
    with pytest.raises(NotImplementedError):
        p[:, 1:3]  # not whole bytes


def test_group_index_select_and_inference_gathers():
    import torch

    from torchrec_b200.ops import jagged as J
    from torchrec_b200.ops.uvm import is_uvm_tensor, new_unified_tensor
    from torchrec_b200.parallel.dist_data import all_to_one_device, merge_pooled_embeddings, sum_reduce_to_one

    g = torch.Generator().manual_seed(0)
    xs = [torch.randn(5, 4, generator=g, requires_grad=True), torch.randn(7, 4, generator=g, requires_grad=True), torch.randn(3, 2, generator=g, requires_grad=True)]
    idx = [torch.tensor([4, 0, 0, 2]), torch.tensor([6, 1]), torch.tensor([2, 2, 1])]
    outs = J.group_index_select_dim0(xs, idx)
    for x, i, o in zip(xs, idx, outs):
        assert torch.equal(o, x.index_select(0, i))
    sum(o.sum() for o in outs).backward()
    assert torch.equal(xs[0].grad[:, 0], torch.tensor([2.0, 0.0, 1.0, 0.0, 1.0]))
    lengths = torch.tensor([1, 2, 0, 1, 3, 1])  # F=3, stride=2
    values = torch.arange(8)
    pl, pv, _ = J.permute_2D_sparse_data_input1D(torch.tensor([2, 0], dtype=torch.int32), lengths, values, 2)
    assert pl.tolist() == [3, 1, 1, 2] and pv.tolist() == [4, 5, 6, 7, 0, 1, 2]
    cpu = torch.device("cpu")
    a, b = torch.ones(2, 3), torch.full((2, 2), 2.0)
    assert merge_pooled_embeddings([a, b], 2, cpu, 1).shape == (2, 5)
    assert torch.equal(sum_reduce_to_one([a, a * 2], cpu), a * 3)
    assert all_to_one_device([a, b], cpu)[1] is b
    u = new_unified_tensor(a, (4, 8))
    assert u.shape == (4, 8) and is_uvm_tensor(u) and not is_uvm_tensor(a)


def test_examples_transfer_learning_and_prediction():
    import importlib.util
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")

    def load(name):
        spec = importlib.util.spec_from_file_location(f"ex_{name}", os.path.join(root, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    assert load("transfer_learning").main() < 0.25
    assert load("prediction").main(steps=120) > 0.7


def _sharding_types(ctx):
    import importlib.util
    import os

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")
    spec = importlib.util.spec_from_file_location("ex_sharding_types", os.path.join(root, "sharding_types.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    res = mod.run(ctx.world_size, ctx.rank, ctx.device)
    assert set(res) == {"table_wise", "row_wise", "column_wise", "data_parallel", "mixed"}
    for shape, _where in res.values():
        assert shape == (8, 2 * 16 * ctx.world_size)


def test_example_sharding_types_gloo():
    from torchrec_b200.utils.multiprocess import run_multi_process

    run_multi_process(_sharding_types, world_size=2, backend="gloo")


def test_virtual_table_eviction_policies_and_feature_scores():
    import torch

    from torchrec_b200.modules.embedding_configs import (CountBasedEvictionPolicy, DataType, EmbeddingBagConfig, FeatureScoreBasedEvictionPolicy, NoEvictionPolicy,
                                                           TimestampBasedEvictionPolicy, eviction_policy_to_cache_algorithm)
    from torchrec_b200.parallel.feature_score_utils import create_sharding_type_to_feature_score_mapping, may_collect_feature_scores
    from torchrec_b200.sparse import KeyedJaggedTensor

    pol = FeatureScoreBasedEvictionPolicy(feature_score_mapping={"fa": 2.0}, feature_score_default_value=0.5)
    a = EmbeddingBagConfig(name="a", embedding_dim=8, num_embeddings=1 << 40, feature_names=["fa", "fa2"], use_virtual_table=True, virtual_table_eviction_policy=pol,
                           data_type=DataType.FP16)
    assert pol.initialized and pol.get_meta_header_len() == 8 and pol.get_embedding_dim() == 8       # 16 B header / 2 B elements
    b = EmbeddingBagConfig(name="b", embedding_dim=8, num_embeddings=100, feature_names=["fb"])
    acc, auto, mapping = create_sharding_type_to_feature_score_mapping([a, b], {"row_wise": ["a", "b"]})
    assert acc and not auto and mapping == {"row_wise": {"fa": 2.0, "fa2": 0.5}}
    assert CountBasedEvictionPolicy(eviction_threshold=3).inference_eviction_threshold == 3 and TimestampBasedEvictionPolicy().inference_eviction_ttl_mins == 1440
    assert eviction_policy_to_cache_algorithm(pol) == "lfu" and eviction_policy_to_cache_algorithm(NoEvictionPolicy()) == "lru"
    kjt = KeyedJaggedTensor(keys=["fa", "fa2"], values=torch.tensor([1, 2, 3]), lengths=torch.tensor([2, 0, 1, 0]))
    scored = may_collect_feature_scores(kjt, True, mapping["row_wise"])
    assert scored.weights().tolist() == [2.0, 2.0, 0.5]
    assert create_sharding_type_to_feature_score_mapping([b], {"table_wise": ["b"]}) == (False, False, {})


def test_criteo_scripts_and_test_utils(tmp_path):
    import numpy as np

    from torchrec_b200.datasets.criteo import CriteoIterDataPipe
    from torchrec_b200.datasets.scripts import contiguous_preproc_criteo, shuffle_preproc_criteo
    from torchrec_b200.datasets.test_utils.criteo_test_utils import CriteoTest

    helper = CriteoTest()
    with CriteoTest._create_dataset_tsv(num_rows=20) as tsv:
        rows = list(CriteoIterDataPipe([tsv]))
        assert len(rows) == 20
        helper._validate_sample(rows[0])
    days = 3
    src = tmp_path / "npy"
    src.mkdir()
    rng = np.random.default_rng(0)
    for d in range(days):
        n = 30 + d
        np.save(src / f"day_{d}_dense.npy", rng.random((n, 13), dtype=np.float32))
        np.save(src / f"day_{d}_sparse.npy", rng.integers(0, 50, size=(n, 26), dtype=np.int32) * 1000003)
        np.save(src / f"day_{d}_labels.npy", np.full((n, 1), d, dtype=np.int32))
    out_c = tmp_path / "contig"
    contiguous_preproc_criteo.main(["--input_dir", str(src), "--output_dir", str(out_c), "--days", str(days), "--frequency_threshold", "2"])
    c0 = np.load(out_c / "day_0_sparse_contig_freq.npy")
    assert c0.shape == (30, 26) and c0.min() >= 1 and c0.max() < 60                       # contiguous: 0 missing, 1 rare, 2.. kept ids
    out_s = tmp_path / "shuf"
    shuffle_preproc_criteo.main(["--input_dir_labels_and_dense", str(src), "--input_dir_sparse", str(src), "--output_dir_shuffled", str(out_s), "--days", str(days)])
    l0, l1 = np.load(out_s / "day_0_labels.npy"), np.load(out_s / "day_1_labels.npy")
    assert l0.shape[0] == 30 and l1.shape[0] == 31 and set(np.unique(np.concatenate([l0, l1])).tolist()) == {0, 1}   # train days mixed, last day untouched
    assert not (out_s / "day_2_labels.npy").exists()
    with CriteoTest._create_dataset_npys(num_rows=12, filenames=["a", "b"]) as paths:
        assert len(paths) == 6 and np.load(paths[1]).shape == (12, 26)


def test_examples_nvt_binary_dataloader_and_data_parallel(tmp_path):
    import importlib.util
    import os

    import numpy as np

    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")

    def load(name):
        spec = importlib.util.spec_from_file_location(f"ex_{name}", os.path.join(root, f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    nvt = load("nvt_dataloader")
    nvt.write_binary_dataset(str(tmp_path), 100, [50] * 26, seed=1)
    raw = np.fromfile(os.path.join(tmp_path, "sparse.bin"), dtype=np.int32).reshape(100, 26)
    ds0, ds1 = nvt.NvtBinaryDataset(str(tmp_path), 16, rank=0, world_size=2), nvt.NvtBinaryDataset(str(tmp_path), 16, rank=1, world_size=2)
    assert len(ds0) == len(ds1) == 3                                   # 6 full batches of 16, strided over 2 ranks
    b = ds1[1]                                                          # global batch 3 -> samples 48..63
    assert b.sparse_features.keys()[3] == "cat_3" and b.sparse_features.stride() == 16
    assert b.sparse_features.values().view(26, 16)[3].tolist() == raw[48:64, 3].tolist()
    assert b.dense_features.shape == (16, 13) and b.dense_features.dtype == torch.float32 and b.labels.shape == (16,)
    assert nvt.main(steps=60, batch_size=128) < 0.97                    # the planted signal is learnt from the files
    assert load("golden_training_data_parallel").main(["--cpu", "--steps", "6", "--batch-size", "64"]) < 1.0


def test_kjt_validator_vbe_and_report():
    import pytest
    import torch

    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.sparse.jagged_tensor import KeyedJaggedTensor
    from torchrec_b200.sparse.jagged_tensor_validator import collect_problems, validate_keyed_jagged_tensor, validate_on_device

    cfgs = [EmbeddingBagConfig(name="t0", embedding_dim=4, num_embeddings=5, feature_names=["a"]),
            EmbeddingBagConfig(name="t1", embedding_dim=4, num_embeddings=3, feature_names=["b"])]
    # variable batch: a has 2 samples, b has 1; full batch 3 through the inverse indices
    vbe = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([0, 1, 2, 2]), lengths=torch.tensor([2, 1, 1]), stride_per_key_per_rank=[[2], [1]],
                            inverse_indices=(["a", "b"], torch.tensor([[0, 1, 0], [0, 0, 0]])))
    assert validate_keyed_jagged_tensor(vbe, cfgs) is True and collect_problems(vbe, cfgs) == []
    bad_inv = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([0, 1, 2, 2]), lengths=torch.tensor([2, 1, 1]), stride_per_key_per_rank=[[2], [1]],
                                inverse_indices=(["a", "b"], torch.tensor([[0, 1, 0], [0, 1, 0]])))  # b has one row: index 1 is outside
    with pytest.raises(ValueError, match="inverse_indices entries"):
        validate_keyed_jagged_tensor(bad_inv)
    oob = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([0, 7, 2, -1]), lengths=torch.tensor([1, 1, 1, 1]))
    rep = collect_problems(oob, cfgs)
    assert len(rep) == 2 and "feature a: 1 of 2" in rep[0] and "feature b: 1 of 2" in rep[1]
    assert int(validate_on_device(oob, cfgs)) == 8 and int(validate_on_device(vbe, cfgs)) == 0
    broken = KeyedJaggedTensor(keys=["a", "b"], values=torch.tensor([0, 1]), offsets=torch.tensor([0, 2, 1, 2, 2]))
    assert int(validate_on_device(broken)) & 2


def test_nvt_preprocessing_chain(tmp_path):
    """TSV day files -> parquet -> hashed / log-scaled parquet -> row-major binary -> per-column binaries; every stage is checked against
    the raw values (missing fields, hex ids above 2^31, odd line count of the last day)."""
    import numpy as np
    import pyarrow.parquet as pq

    from torchrec_b200.datasets.scripts.nvt import convert_parquet_to_binary, convert_tsv_to_parquet, process_criteo_parquet, split_binary_dataset
    from torchrec_b200.datasets.scripts.nvt.utils.criteo_constant import DEFAULT_CAT_NAMES, DEFAULT_INT_NAMES, NUM_EMBEDDINGS_PER_FEATURE

    rng = np.random.default_rng(0)
    raw, days = {}, 3
    tsv = tmp_path / "tsv"
    tsv.mkdir()
    for d in range(days):
        n = 21 if d == days - 1 else 16
        rows = []
        for _ in range(n):
            label = int(rng.integers(0, 2))
            ints = [None if rng.random() < 0.2 else int(rng.integers(0, 1000)) for _ in range(13)]
            cats = [None if rng.random() < 0.2 else int(rng.integers(0, 2**32)) for _ in range(26)]
            rows.append((label, ints, cats))
        raw[d] = rows
        with open(tsv / f"day_{d}", "w") as f:
            for label, ints, cats in rows:
                f.write("\t".join([str(label)] + ["" if x is None else str(x) for x in ints] + ["" if c is None else format(c, "x") for c in cats]) + "\n")
    base = tmp_path / "out"
    files = convert_tsv_to_parquet.convert_tsv_to_parquet(str(tsv), str(base), days=days)
    assert [os.path.basename(f) for f in files] == ["day_0.parquet", "day_1.parquet", "day_2.part0.parquet", "day_2.part1.parquet"]
    t0 = pq.read_table(files[0]).to_pydict()
    assert t0["label"] == [r[0] for r in raw[0]] and t0["int_3"] == [r[1][3] for r in raw[0]] and t0["cat_7"] == [r[2][7] for r in raw[0]]
    part0, part1 = pq.read_table(files[2]), pq.read_table(files[3])
    assert part0.num_rows == 11 and part1.num_rows == 10 and part0.to_pydict()["cat_0"] + part1.to_pydict()["cat_0"] == [r[2][0] for r in raw[2]]
    # hashing + log scaling
    out = process_criteo_parquet.process(str(base), shuffle_train=False, days=days)
    tr = pq.read_table(os.path.join(out, "train", "part_1.parquet")).to_pydict()
    for i, (label, ints, cats) in enumerate(raw[1]):
        assert tr["label"][i] == float(label)
        assert tr["int_5"][i] == pytest.approx(0.0 if ints[5] is None else float(np.log(ints[5] + 3.0)), rel=1e-6)
        want = 0 if cats[2] is None else int(process_criteo_parquet.hash_bucket(np.array([cats[2]]), NUM_EMBEDDINGS_PER_FEATURE[2])[0])
        assert tr["cat_2"][i] == want and 0 <= tr["cat_2"][i] < NUM_EMBEDDINGS_PER_FEATURE[2]
    shuffled = process_criteo_parquet.process(str(base), shuffle_train=True, days=days)
    sh = pq.read_table(os.path.join(shuffled, "train", "part_0.parquet")).to_pydict()
    plain = pq.read_table(os.path.join(out, "train", "part_0.parquet"))
    assert sorted(sh["cat_9"]) == sorted(plain.to_pydict()["cat_9"]) and pq.read_table(os.path.join(shuffled, "validation", "part_0.parquet")).num_rows == 11
    # binaries
    convert_parquet_to_binary.convert(shuffled, str(tmp_path / "inter"), str(tmp_path / "bin"))
    rec = np.fromfile(tmp_path / "bin" / "train_data.bin", dtype=convert_parquet_to_binary.RECORD_DTYPE)
    assert rec.shape == (32,) and rec["cat_9"][:16].tolist() == sh["cat_9"] and rec["int_0"][:16].tolist() == pytest.approx(sh["int_0"])
    split_binary_dataset.split_dataset(str(tmp_path / "bin"), str(tmp_path / "split"), batch_size=5)
    num = np.fromfile(tmp_path / "split" / "train" / "numerical.bin", dtype=np.float32).reshape(-1, 13)
    lab = np.fromfile(tmp_path / "split" / "train" / "label.bin", dtype=np.float32)
    c4 = np.fromfile(tmp_path / "split" / "train" / "cat_4.bin", dtype=np.int32)
    assert num.shape == (32, 13) and np.array_equal(num[:, 0], rec["int_0"]) and np.array_equal(lab, rec["label"].astype(np.float32)) and np.array_equal(c4, rec["cat_4"])
    assert np.fromfile(tmp_path / "split" / "test" / "label.bin", dtype=np.float32).shape == (10,)
    assert set(DEFAULT_INT_NAMES) | set(DEFAULT_CAT_NAMES) <= set(plain.column_names)
    # the dataloader of the example reads that layout: rank-strided whole batches
    import importlib.util

    spec = importlib.util.spec_from_file_location("nvt_dataloader", os.path.join(os.path.dirname(__file__), "..", "examples", "nvt_dataloader.py"))
    nvt = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nvt)
    ds = nvt.NvtSplitBinaryDataset(str(tmp_path / "split" / "train"), batch_size=8, rank=1, world_size=2)
    assert len(ds) == 2
    b1 = ds[1]  # global batch 3: rows 24..31
    assert torch.equal(b1.labels, torch.from_numpy(lab[24:32])) and torch.equal(b1.dense_features, torch.from_numpy(num[24:32]))
    assert b1.sparse_features["cat_4"].values().tolist() == c4[24:32].tolist() and b1.sparse_features.stride() == 8
    # command lines
    convert_tsv_to_parquet.main(["-i", str(tsv), "-o", str(tmp_path / "cli"), "--days", str(days)])
    process_criteo_parquet.main(["-b", str(tmp_path / "cli"), "--days", str(days)])
    convert_parquet_to_binary.main(["--src_dir", str(tmp_path / "cli" / "criteo_preproc"), "--intermediate_dir", str(tmp_path / "cli" / "i"), "--dst_dir", str(tmp_path / "cli" / "b"),
                                    "--parallel_jobs", "2"])
    split_binary_dataset.main(["--input_path", str(tmp_path / "cli" / "b"), "--output_path", str(tmp_path / "cli" / "s"), "--batch_size", "7"])
    assert np.array_equal(np.fromfile(tmp_path / "cli" / "s" / "validation" / "cat_0.bin", dtype=np.int32),
                          np.fromfile(tmp_path / "split" / "validation" / "cat_0.bin", dtype=np.int32))
