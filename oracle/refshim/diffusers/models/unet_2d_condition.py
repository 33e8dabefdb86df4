class UNet2DConditionModel:  # only referenced in a type annotation (sparse_controlnet.py:38)
    pass
