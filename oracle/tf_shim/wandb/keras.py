class WandbCallback:                           # utils.py:5 imports it; never constructed on the model path
    def __init__(self, *a, **k):
        raise RuntimeError("wandb is not available in the oracle shim")
