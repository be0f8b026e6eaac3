"""Batch-1 serving forward: graph replay time and phase stamps, shared-memory exchange vs global-scratch exchange."""
import sys, json, torch
sys.path.insert(0, "/root/repo")
import mmer_b200 as mm
from mmer_b200 import _lib
import bench

dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
m = mm.MultimodalEmotionModel(max_seq_len=6, fusion_num_layers=2, classifier_hidden_dim=512).to(dev).eval()
m.compute_dtype = torch.bfloat16
v, a = torch.randn(1, 5, 768, device=dev).bfloat16(), torch.randn(1, 1024, device=dev).bfloat16()
mk = torch.zeros(1, 5, dtype=torch.bool, device=dev)
out = {}
for name, knob in (("small", 0), ("global", 1), ("shared", 2)):
    lib.mmer_debug_set(_lib.DEBUG_SERVE_GLOBAL, knob)
    srv = mm.ServingForward(m, frames=5)
    srv(v, a, mk)
    out[name] = {"call_us": bench._event_ms(lambda: srv(v, a, mk), 300, warm=20) * 1e3,
                 "graph_only_us": bench._event_ms(srv.graph.replay, 300, warm=20) * 1e3,
                 "phase_ns": srv.phase_times()}
    if isinstance(out[name]["phase_ns"], (list, tuple)):
        out[name]["kernel_ns"] = sum(out[name]["phase_ns"])
lib.mmer_debug_set(_lib.DEBUG_SERVE_GLOBAL, 0)
print(json.dumps(out))
lib.mmer_debug_set(_lib.DEBUG_SERVE_STAMPS, 1)
srv = mm.ServingForward(m, frames=5)
for _ in range(5):
    srv(v, a, mk)
print(json.dumps({"small_fine_stamps_ns": srv.phase_times()}))
lib.mmer_debug_set(_lib.DEBUG_SERVE_STAMPS, 0)
