"""Which descriptor form the row-shifted activation operand needs (run once on the GPU): error of each `halo` mode against the goldens."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from usdm_b200.vocoder import BigVGAN
from oracle import vocoder_oracle as V
for name, cfg in (("vocoder_small_b2.pt", V.SMALL_CFG), ("vocoder_full_t6.pt", V.BIGVGAN_22K_80)):
    g = torch.load(os.path.join("tests", "golden", name), weights_only=False)
    m = BigVGAN(cfg)
    m.load_state_dict(V.init_vocoder_weights(g["seed"], cfg))
    m = m.cuda().eval()
    for halo in (0, 1):
        m.set_option("halo", halo)
        wav = m(g["mel"].cuda()).cpu()
        print(name, "halo", halo, "rel-L2", float((wav - g["wav"]).norm() / g["wav"].norm()), flush=True)
