def scale_lora_layers(model, weight):
    return None


def unscale_lora_layers(model, weight=None):
    return None
