import os, sys, time, numpy as np, torch
sys.path.insert(0, os.getcwd())
from pmgt_b200 import trainer
dev = torch.device("cuda", 0)
args = trainer.make_args(synthetic="TG", train_batch_size=4096, seed=0); args.device = dev
trainer.set_seed(0)
args.graph, args.feat_init_emb = trainer._load_graph_and_features(args)
trainer.init_dataloader(args); trainer.init_model(args)
tm = trainer.PMGTTrainerModel(args); ds = args.train_dataset
idx = [torch.from_numpy(np.resize(trainer.epoch_permutation(len(ds), 0, s), 4096).astype(np.int64)).to(dev) for s in range(13)]
for s in range(3): tm.train_on_indices(ds, idx[s], epoch=s)
torch.cuda.synchronize()
t0 = time.perf_counter()
for s in range(3, 13): tm.train_on_indices(ds, idx[s], epoch=s)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host issue {1e2*(t1-t0):.2f} ms/step, total {1e2*(t2-t0):.2f} ms/step")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for s in range(3, 8): tm.train_on_indices(ds, idx[s], epoch=s)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
