"""comfy.utils: common_upscale (non-bislerp modes are F.interpolate), repeat_to_batch_size."""
import math

import torch


def common_upscale(samples, width, height, upscale_method, crop):
    if crop == "center":
        old_width, old_height = samples.shape[-1], samples.shape[-2]
        old_aspect, new_aspect = old_width / old_height, width / height
        x = y = 0
        if old_aspect > new_aspect:
            x = round((old_width - old_width * (new_aspect / old_aspect)) / 2)
        elif old_aspect < new_aspect:
            y = round((old_height - old_height * (old_aspect / new_aspect)) / 2)
        samples = samples.narrow(-2, y, old_height - y * 2).narrow(-1, x, old_width - x * 2)
    if upscale_method == "bislerp":
        raise NotImplementedError("bislerp is implemented inside ComfyUI and is not restated here")
    return torch.nn.functional.interpolate(samples, size=(height, width), mode=upscale_method)


def repeat_to_batch_size(tensor, batch_size, dim=0):
    if tensor.shape[dim] > batch_size:
        return tensor.narrow(dim, 0, batch_size)
    if tensor.shape[dim] < batch_size:
        reps = dim * [1] + [math.ceil(batch_size / tensor.shape[dim])] + [1] * (len(tensor.shape) - 1 - dim)
        return tensor.repeat(reps).narrow(dim, 0, batch_size)
    return tensor
