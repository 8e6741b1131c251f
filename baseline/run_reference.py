"""Reference arm of bench.py: the UNMODIFIED reference (baseline/_ref) on the same metric / config, through its own public
API (EmbeddingBagCollection -> DLRM -> DistributedModelParallel -> TrainPipelineSparseDist) and stock fbgemm kernels.
Nothing from torchrec_b200 is imported here. Only reachable when `import torchrec` works (needs fbgemm_gpu)."""
import json
import os
import time


def run(args) -> None:
    import torch
    import torch.distributed as dist
    from torchrec import EmbeddingBagCollection
    from torchrec.datasets.random import RandomRecDataset
    from torchrec.distributed import DistributedModelParallel, TrainPipelineSparseDist
    from torchrec.distributed.embeddingbag import EmbeddingBagCollectionSharder
    from torchrec.distributed.planner import EmbeddingShardingPlanner, Topology
    from torchrec.distributed.planner.types import ParameterConstraints
    from torchrec.models.dlrm import DLRM, DLRMTrain
    from torchrec.modules.embedding_configs import EmbeddingBagConfig
    from torchrec.optim.apply_optimizer_in_backward import apply_optimizer_in_backward
    from torchrec.optim.keyed import CombinedOptimizer, KeyedOptimizerWrapper
    from torchrec.optim.optimizers import in_backward_optimizer_filter
    from torchrec.optim.rowwise_adagrad import RowWiseAdagrad

    from bench import CRITEO_1TB_40M, ClockSampler  # config constants + clock sampler only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    if not dist.is_initialized():
        dist.init_process_group(backend="nccl", rank=rank, world_size=world)
    hashes = [min(h, args.row_cap) for h in CRITEO_1TB_40M]
    keys = [f"cat_{i}" for i in range(26)]
    tables = [EmbeddingBagConfig(name=f"t_{k}", embedding_dim=args.embedding_dim, num_embeddings=h, feature_names=[k]) for k, h in zip(keys, hashes)]
    ebc = EmbeddingBagCollection(tables=tables, device=torch.device("meta"))
    apply_optimizer_in_backward(RowWiseAdagrad, ebc.parameters(), {"lr": args.lr, "eps": 1e-8})
    model = DLRMTrain(DLRM(embedding_bag_collection=ebc, dense_in_features=13, dense_arch_layer_sizes=[int(x) for x in args.dense_arch.split(",")],
                           over_arch_layer_sizes=[int(x) for x in args.over_arch.split(",")], dense_device=device))
    planner = EmbeddingShardingPlanner(topology=Topology(world_size=world, local_world_size=world, compute_device="cuda"), batch_size=args.batch_size,
                                       constraints={t.name: ParameterConstraints(sharding_types=["table_wise"], compute_kernels=["fused"]) for t in tables})
    sharders = [EmbeddingBagCollectionSharder()]
    plan = planner.collective_plan(model, sharders, dist.GroupMember.WORLD)
    dmp = DistributedModelParallel(model, device=device, plan=plan, sharders=sharders)
    dense_opt = KeyedOptimizerWrapper(dict(in_backward_optimizer_filter(dmp.named_parameters())), lambda p: torch.optim.SGD(p, lr=args.lr))
    opt = CombinedOptimizer([dmp.fused_optimizer, dense_opt])
    B = args.batch_size
    ds = RandomRecDataset(keys, B, hash_sizes=hashes, ids_per_features=[args.pooling] * 26, num_dense=13, manual_seed=1234 + rank,
                          num_generated_batches=args.num_host_batches)
    host = [b.pin_memory() for b in ds.batch_generator._generated_batches]
    W = max(args.warmup, 3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # ---- `value`: the plain training step on device-resident batches (the same measurement as the other arm's `value`) ----
    dev_batches = [b.to(device) for b in host]

    def step(b):
        opt.zero_grad()
        loss, _ = dmp(b)
        loss.backward()
        opt.step()

    for i in range(W):
        step(dev_batches[i % len(dev_batches)])
    dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.mark_start()
    e0.record()
    for i in range(args.steps):
        step(dev_batches[i % len(dev_batches)])
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    tv = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    ms_value = float(tv.item())
    clocks = sampler.stop()

    # ---- `e2e`: TrainPipelineSparseDist fed from pinned host batches, loss read back every step ----
    pipe = TrainPipelineSparseDist(dmp, opt, device)

    def it(n):
        for i in range(n):
            yield host[i % len(host)]

    stream = it(W + args.steps + 2)
    for _ in range(W):
        pipe.progress(stream)
    dist.barrier()
    torch.cuda.synchronize()
    loss_host = torch.zeros(1).pin_memory()
    e0.record()
    for _ in range(args.steps):
        out = pipe.progress(stream)
        loss_host.copy_(out[0].detach().reshape(1), non_blocking=True)
    e1.record()
    dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        v = B * world * args.steps / (ms_value / 1e3)
        v2 = B * world * args.steps / (ms / 1e3)
        b0 = host[0]
        kjt = b0.sparse_features
        h2d = sum(x.numel() * x.element_size() for x in (b0.dense_features, b0.labels, kjt.values(), kjt.lengths()) if x is not None)
        print(json.dumps({"metric": "DLRM training throughput (samples/s, whole job, device-timed, max over ranks)", "value": v, "unit": "samples/s", "n_gpus": world,
                          "steps": args.steps, "warmup": W, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "fp32", "data": "synthetic", "impl": "reference", "clocks": clocks,
                          "e2e": {"value": v2, "unit": "samples/s", "ms_per_step": ms / args.steps, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
                          "config": {"model": "reference torchrec DLRM, same tables / arch / batch", "global_batch": B * world,
                                     "parallelism": "table_wise (EmbeddingShardingPlanner, same constraint as the other arm) + DDP"}}))
    dist.destroy_process_group()
