"""BERT4Rec-style sequential recommender (reference examples/bert4rec): an (unpooled) EmbeddingCollection feeds a
Transformer encoder; masked-item prediction. The item table is the model-parallel part (sharded by
EmbeddingCollectionSharder under torchrun), the encoder is data parallel."""
import os
import sys

import torch
from torch import nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from torchrec_b200.modules.embedding_configs import EmbeddingConfig  # noqa: E402
from torchrec_b200.modules.embedding_modules import EmbeddingCollection  # noqa: E402
from torchrec_b200.sparse import KeyedJaggedTensor  # noqa: E402


class BERT4Rec(nn.Module):
    """``forward(KJT of item histories) -> logits [B, L, vocab]`` (padded to ``max_len``)."""

    def __init__(self, vocab_size: int, max_len: int, emb_dim: int = 64, nhead: int = 2, num_layers: int = 2, device=None) -> None:
        super().__init__()
        self.max_len = max_len
        self.item_embedding = EmbeddingCollection([EmbeddingConfig(name="item_embedding", embedding_dim=emb_dim, num_embeddings=vocab_size + 2, feature_names=["item"])], device=device)
        self.position = nn.Embedding(max_len, emb_dim, device=device)
        layer = nn.TransformerEncoderLayer(emb_dim, nhead, emb_dim * 4, dropout=0.1, batch_first=True, device=device)
        self.encoder = nn.TransformerEncoder(layer, num_layers)
        self.out = nn.Linear(emb_dim, vocab_size + 2, device=device)

    def forward(self, history: KeyedJaggedTensor) -> torch.Tensor:
        jt = self.item_embedding(history)["item"]
        x = jt.to_padded_dense(self.max_len)  # [B, L, D]
        lengths = jt.lengths()
        mask = torch.arange(self.max_len, device=x.device).unsqueeze(0) >= lengths.unsqueeze(1)
        x = x + self.position.weight.unsqueeze(0)
        return self.out(self.encoder(x, src_key_padding_mask=mask))


def main() -> None:
    torch.manual_seed(0)
    vocab, L, B = 500, 16, 32
    model = BERT4Rec(vocab, L)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    g = torch.Generator().manual_seed(0)
    MASK = vocab + 1
    for step in range(8):
        lengths = torch.randint(4, L + 1, (B,), generator=g)
        items = torch.randint(1, vocab + 1, (int(lengths.sum()),), generator=g)
        masked = items.clone()
        pick = torch.rand(items.numel(), generator=g) < 0.2
        masked[pick] = MASK
        logits = model(KeyedJaggedTensor(keys=["item"], values=masked, lengths=lengths))
        # scatter the labels of the masked positions into the padded layout
        offs = torch.cat([lengths.new_zeros(1), lengths.cumsum(0)])
        row = torch.repeat_interleave(torch.arange(B), lengths)
        col = torch.arange(items.numel()) - offs[row]
        tgt = torch.full((B, L), -100, dtype=torch.long)
        tgt[row[pick], col[pick]] = items[pick]
        loss = nn.functional.cross_entropy(logits.reshape(-1, vocab + 2), tgt.reshape(-1), ignore_index=-100)
        opt.zero_grad()
        loss.backward()
        opt.step()
        if step % 2 == 0:
            print(f"step {step}: masked-item loss {float(loss):.4f}")


if __name__ == "__main__":
    main()
