"""Request construction for the ``Predictor`` gRPC service (reference ``inference/client.py:20-121``): a training ``Batch`` becomes a
``PredictionRequest`` whose ``float_features`` / ``id_list_features`` carry raw little-endian bytes (fp32 values, int32 lengths,
int32/int64 ids), exactly what ``predictor.proto`` defines. ``python -m torchrec_b200.inference.client --ip HOST --port 50051``."""
from __future__ import annotations

import argparse
from typing import Any, List

import numpy as np
import torch

from ..datasets.utils import Batch
from .dlrm_predict import create_training_batch
from .server import PredictorClient, proto_classes


def create_request(batch: Batch, num_dense: int, num_id_list_features: int) -> Any:
    """``PredictionRequest`` protobuf for ``batch`` (built with the runtime-generated message classes of ``server.py``)."""
    Request, _Response = proto_classes()

    def to_bytes(t: torch.Tensor) -> bytes:
        return t.detach().cpu().contiguous().numpy().tobytes()

    kjt = batch.sparse_features
    req = Request(batch_size=batch.dense_features.shape[0])
    req.float_features.num_features = num_dense
    req.float_features.values = to_bytes(batch.dense_features.float())
    req.id_list_features.num_features = num_id_list_features
    req.id_list_features.values = to_bytes(kjt.values().to(torch.int32))
    req.id_list_features.lengths = to_bytes(kjt.lengths().to(torch.int32))
    return req


def main(argv: List[str] = None) -> None:  # type: ignore[assignment]
    ap = argparse.ArgumentParser()
    ap.add_argument("--ip", type=str, default="127.0.0.1")
    ap.add_argument("--port", type=int, default=50051)
    ap.add_argument("--num_float_features", type=int, default=13)
    ap.add_argument("--num_id_list_features", type=int, default=26)
    ap.add_argument("--num_embeddings", type=int, default=100000)
    ap.add_argument("--batch_size", type=int, default=100)
    a = ap.parse_args(argv)
    keys = [f"cat_{i}" for i in range(a.num_id_list_features)]
    batch = create_training_batch(a.num_float_features, keys, a.num_embeddings, a.batch_size)
    client = PredictorClient(f"{a.ip}:{a.port}")
    kjt = batch.sparse_features
    out = client.predict(a.batch_size, dense=batch.dense_features.numpy().astype(np.float32),
                         id_list=(kjt.lengths().numpy().astype(np.int32), kjt.values().numpy().astype(np.int32)), num_id_list_features=a.num_id_list_features)
    print("predictions:", out[:10])
    client.close()


if __name__ == "__main__":
    main()
