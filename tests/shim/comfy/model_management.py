import torch


def get_torch_device():
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def device_supports_non_blocking(device):
    return torch.device(device).type == "cuda"


def throw_exception_if_processing_interrupted():
    return None
