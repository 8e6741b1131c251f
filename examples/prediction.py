"""Train-and-predict on a MovieLens-shaped task (reference examples/prediction/predict_using_torchrec.py): user / movie id
features -> EmbeddingBagCollection -> MLP -> rating probability; then score candidate movies for one user.

    python examples/prediction.py"""
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.modules.embedding_configs import EmbeddingBagConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingBagCollection  # noqa: E402
from torchrec_b200.modules.mlp import MLP  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402


class RatingModel(nn.Module):
    def __init__(self, num_users: int, num_movies: int, dim: int = 16) -> None:
        super().__init__()
        self.ebc = EmbeddingBagCollection([EmbeddingBagConfig(name="t_user", embedding_dim=dim, num_embeddings=num_users, feature_names=["user_id"]),
                                           EmbeddingBagConfig(name="t_movie", embedding_dim=dim, num_embeddings=num_movies, feature_names=["movie_id"])])
        self.over = MLP(2 * dim, [32, 1], activation=torch.relu)

    def forward(self, kjt: KeyedJaggedTensor) -> torch.Tensor:
        kt = self.ebc(kjt)
        d = kt.to_dict()
        # matrix-factorisation term (user . movie) plus an MLP over the concatenated embeddings
        return (d["user_id"] * d["movie_id"]).sum(1) + self.over(kt.values()).squeeze(1)


def _kjt(users: torch.Tensor, movies: torch.Tensor) -> KeyedJaggedTensor:
    n = users.numel()
    return KeyedJaggedTensor(keys=["user_id", "movie_id"], values=torch.cat([users, movies]), lengths=torch.ones(2 * n, dtype=torch.int64))


def main(steps: int = 150) -> float:
    torch.manual_seed(0)
    U, M = 200, 300
    taste, genre = torch.randn(U, 4), torch.randn(M, 4)  # latent ground truth: a user likes a movie when the dot product is positive
    model = RatingModel(U, M)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    g = torch.Generator().manual_seed(1)
    for _ in range(steps):
        u, m = torch.randint(0, U, (256,), generator=g), torch.randint(0, M, (256,), generator=g)
        y = ((taste[u] * genre[m]).sum(1) > 0).float()
        loss = nn.functional.binary_cross_entropy_with_logits(model(_kjt(u, m)), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
    u, m = torch.randint(0, U, (2000,), generator=g), torch.randint(0, M, (2000,), generator=g)
    with torch.no_grad():
        acc = float(((model(_kjt(u, m)) > 0) == ((taste[u] * genre[m]).sum(1) > 0)).float().mean())
        scores = torch.sigmoid(model(_kjt(torch.full((M,), 7), torch.arange(M))))
    print(f"held-out accuracy {acc:.3f}; top-5 movies for user 7: {scores.topk(5).indices.tolist()}")
    return acc


if __name__ == "__main__":
    main()
