from .asg import ASGLoss  # noqa: F401  (same export as the reference's torch_asg/__init__.py:1)
from .graphed import GraphedASGLoss  # noqa: F401  (not in the reference: ASGLoss with a workspace + CUDA graphs behind it)
from .peer import PeerAllReduce  # noqa: F401  (not in the reference: the batch-sharded exchange over NVLink peer memory)
from .logits import ASGLossFromLogits  # noqa: F401  (not in the reference: log_softmax fused into the criterion)
