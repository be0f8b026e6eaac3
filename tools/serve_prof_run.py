"""One batch-1 serving request (serve_small.cu) for an ncu capture: python tools/serve_prof_run.py"""
import sys, torch
sys.path.insert(0, "/root/repo")
import mmer_b200 as mm

dev = torch.device("cuda:0")
torch.manual_seed(0)
m = mm.MultimodalEmotionModel(max_seq_len=6, fusion_num_layers=2, classifier_hidden_dim=512).to(dev).eval()
m.compute_dtype = torch.bfloat16
v, a = torch.randn(1, 5, 768, device=dev).bfloat16(), torch.randn(1, 1024, device=dev).bfloat16()
mk = torch.zeros(1, 5, dtype=torch.bool, device=dev)
srv = mm.ServingForward(m, frames=5, use_graph=False)
for _ in range(4):
    srv(v, a, mk)
torch.cuda.synchronize()
print("ok")
