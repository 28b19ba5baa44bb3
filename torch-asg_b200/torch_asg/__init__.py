from .asg import ASGLoss  # noqa: F401  (same export as the reference's torch_asg/__init__.py:1)
