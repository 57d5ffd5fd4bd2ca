import os, sys, traceback
sys.path.insert(0, '/root/repo')
import torch
from lsnet_b200.data import MODEL_CFG, synthetic_batch
from lsnet_b200.train import GraphTrainer
b = synthetic_batch(0, batch=1, img_hw=(256, 320), pin=True)
try:
    tr = GraphTrainer(MODEL_CFG['bbox_r50'], b, kernel_timing=True)
    print('capture ok')
except Exception as e:
    traceback.print_exc()
