import os, sys, torch
sys.path.insert(0, "/root/repo")
os.environ["B200VQA_ENC_FFN_DBG"] = "1"
os.environ["B200VQA_NO_GRAPH"] = "1"
from explainable_spatial_vqa_b200 import inference_transformer_iqap as iq
from explainable_spatial_vqa_b200 import synthetic as syn
torch.manual_seed(0)
m = iq.VQAModel(85, 256, 256, 32, 44, 27, 196).eval().cuda()
img, q = syn.iqap_inputs(1024, seed=1)
img, q = img.cuda(), q.cuda()
for _ in range(3):
    m(img, q)
torch.cuda.synchronize()
