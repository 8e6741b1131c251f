"""Native inference server core: batching, result split, back-pressure, gRPC front
(reference inference_legacy/tests/BatchingQueueTest.cpp, BatchingTest.cpp, ResultSplitTest.cpp)."""
import threading

import numpy as np
import pytest
import torch


def _model(keys):
    from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig
    from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection

    torch.manual_seed(0)
    ebc = EmbeddingBagCollection([EmbeddingBagConfig(name=f"t_{k}", embedding_dim=4, num_embeddings=50, feature_names=[k]) for k in keys])
    lin = torch.nn.Linear(3 + 4 * len(keys), 1)

    def predict(dense, kjt, wkjt):
        return torch.sigmoid(lin(torch.cat([dense, ebc(kjt).values()], 1))).squeeze(1)

    return predict


def _request(rng, B, F):
    dense = rng.standard_normal((B, 3)).astype(np.float32)
    lengths = rng.integers(0, 4, size=F * B).astype(np.int32)
    values = rng.integers(0, 50, size=int(lengths.sum())).astype(np.int64)
    return dense, lengths, values


def _reference(predict, keys, dense, lengths, values, B):
    from torchrec_b200.sparse import KeyedJaggedTensor

    with torch.no_grad():
        return predict(torch.from_numpy(dense), KeyedJaggedTensor(keys=keys, values=torch.from_numpy(values), lengths=torch.from_numpy(lengths), stride=B), None).numpy()


def test_batching_and_result_split():
    from torchrec_b200.inference.server import InferenceServer, ServerConfig

    keys = ["a", "b"]
    predict = _model(keys)
    seen_batches = []

    def spy(dense, kjt, wkjt):
        seen_batches.append(dense.shape[0])
        return predict(dense, kjt, wkjt)

    srv = InferenceServer([spy], [torch.device("cpu")], id_list_keys=keys, config=ServerConfig(max_batch_size=64, batching_interval_ms=30.0))
    rng = np.random.default_rng(0)
    reqs = [_request(rng, B, len(keys)) for B in (3, 5, 1, 7, 2, 4)]
    results = [None] * len(reqs)

    def client(i):
        d, l, v = reqs[i]
        results[i] = srv.predict(d.shape[0], d, (l, v))

    ts = [threading.Thread(target=client, args=(i,)) for i in range(len(reqs))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    for (d, l, v), out in zip(reqs, results):
        np.testing.assert_allclose(out[:, 0], _reference(predict, keys, d, l, v, d.shape[0]), rtol=1e-5, atol=1e-6)
    st = srv.stats()
    assert st["requests"] == 6 and st["samples"] == 22
    assert st["batches"] < 6 and max(seen_batches) > 7, (st, seen_batches)  # requests were coalesced
    srv.shutdown()


def test_max_batch_size_and_errors():
    from torchrec_b200.inference.server import InferenceServer, ServerConfig

    keys = ["a"]
    predict = _model(keys)
    sizes = []

    def spy(dense, kjt, wkjt):
        sizes.append(dense.shape[0])
        if dense.shape[0] == 13:
            raise ValueError("boom")
        return predict(dense, kjt, wkjt)

    srv = InferenceServer([spy], [torch.device("cpu")], id_list_keys=keys, config=ServerConfig(max_batch_size=8, batching_interval_ms=20.0))
    rng = np.random.default_rng(1)
    rids = []
    reqs = [_request(rng, 4, 1) for _ in range(5)]
    for d, l, v in reqs:
        rids.append(srv.submit(4, d, (l, v)))
    outs = [srv.wait(r, 4) for r in rids]
    assert max(sizes) <= 8
    for (d, l, v), o in zip(reqs, outs):
        np.testing.assert_allclose(o[:, 0], _reference(predict, keys, d, l, v, 4), rtol=1e-5, atol=1e-6)
    d, l, v = _request(rng, 13, 1)  # oversize request forms its own batch; its failure is reported, the server survives
    with pytest.raises(RuntimeError):
        srv.predict(13, d, (l, v))
    d, l, v = _request(rng, 2, 1)
    assert srv.predict(2, d, (l, v)).shape == (2, 1)
    srv.shutdown()


def test_grpc_front_end():
    grpc = pytest.importorskip("grpc")
    from torchrec_b200.inference.server import InferenceServer, PredictorClient, ServerConfig, serve_grpc

    keys = ["a", "b"]
    predict = _model(keys)
    srv = InferenceServer([predict], [torch.device("cpu")], id_list_keys=keys, config=ServerConfig(max_batch_size=32, batching_interval_ms=2.0))
    g = serve_grpc(srv, port=0)
    client = PredictorClient(f"127.0.0.1:{g.bound_port}")
    rng = np.random.default_rng(2)
    d, l, v = _request(rng, 6, 2)
    out = client.predict(6, d, (l, v), num_id_list_features=2)
    np.testing.assert_allclose(out["default"], _reference(predict, keys, d, l, v, 6), rtol=1e-5, atol=1e-6)
    client.close()
    g.stop(0)
    srv.shutdown()


def test_native_network_front_concurrent_clients_and_malformed_input():
    """C++ TCP front (csrc/serving/net_front.cpp): protobuf requests built by the Python protobuf runtime are parsed by the hand-written
    wire reader, batched with the requests of other connections, and the hand-encoded responses parse back with the protobuf runtime."""
    from torchrec_b200.inference.server import InferenceServer, NativePredictorClient, ServerConfig, serve_native

    keys = ["a", "b"]
    predict = _model(keys)
    srv = InferenceServer([predict], [torch.device("cpu")], id_list_keys=keys, config=ServerConfig(max_batch_size=64, batching_interval_ms=2.0))
    front = serve_native(srv, port=0, task_name="ctr")
    try:
        assert front.port > 0
        errors, done = [], []

        def client(seed: int) -> None:
            try:
                rng = np.random.default_rng(seed)
                c = NativePredictorClient("127.0.0.1", front.port)
                for _ in range(5):
                    B = int(rng.integers(1, 9))
                    dense, lengths, values = _request(rng, B, len(keys))
                    out = c.predict(B, dense=dense, id_list=(lengths, values), num_id_list_features=len(keys))
                    np.testing.assert_allclose(out["ctr"], _reference(predict, keys, dense, lengths, values, B), rtol=1e-5, atol=1e-6)
                c.close()
                done.append(seed)
            except Exception as e:  # pragma: no cover
                errors.append(repr(e))

        threads = [threading.Thread(target=client, args=(s,)) for s in range(6)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(timeout=30)
        assert not errors and len(done) == 6, errors
        st = srv.stats()
        assert st["requests"] == 30 and st["batches"] < 30, "requests of different connections share batches"
        # a malformed frame gets an error response (status field, empty predictions) and the connection stays usable
        c = NativePredictorClient("127.0.0.1", front.port)
        raw = c.predict_raw(b"\x08\x02\x12\x05\x08\x03\x12\x01\x00")  # batch 2, 3 float features, but 1 byte of values
        assert raw == b"\x78\x03"
        dense, lengths, values = _request(np.random.default_rng(99), 2, len(keys))
        assert c.predict(2, dense=dense, id_list=(lengths, values), num_id_list_features=len(keys))["ctr"].shape == (2,)
        c.close()
        fs = front.stats()
        assert fs["served"] == 31 and fs["malformed"] == 1
    finally:
        front.stop()
        srv.shutdown()
