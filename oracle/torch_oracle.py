"""TEST INFRASTRUCTURE (oracle): the policy-net graph as torch CPU ops (oneDNN conv2d) -- the
"torch-CPU fp32 restatement of the identical graph" SURVEY 8(d) names as the CPU stand-in for the
reference's TF-1.0 session, with intra-op threads = cpu_count like the reference's session config
(policy_net_utils.pyx:70-72).  Used by tests and by bench.py's cpu_baseline leg only."""
import os

import numpy as np


def forward(planes, weights, threads=None):
    """planes [N,8,8,4] -> (logits, probs) [N,155] float32; same graph as oracle.forward_fp32
    (policy_net_utils.pyx:106-124,149-178,215-239)."""
    import torch
    import torch.nn.functional as F
    torch.set_num_threads(threads or os.cpu_count() or 1)
    with torch.no_grad():
        x = torch.from_numpy(np.asarray(planes).astype(np.float32)).permute(0, 3, 1, 2)      # NHWC -> NCHW
        for i in range(1, 6):
            k = torch.from_numpy(weights["hidden_layer/%d/weights" % i]).permute(3, 2, 0, 1).contiguous()   # HWIO -> OIHW
            x = F.relu(F.conv2d(x, k, torch.from_numpy(weights["hidden_layer/%d/bias" % i]),
                                padding=(k.shape[-1] - 1) // 2))
        h = x.permute(0, 2, 3, 1).reshape(-1, 64)                                            # row-major r*8+c
        logits = h @ torch.from_numpy(weights["output_layer/weights"]) + torch.from_numpy(weights["output_layer/bias"])
        return logits.numpy(), torch.softmax(logits, dim=1).numpy()
