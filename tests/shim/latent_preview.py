"""Stand-in for ComfyUI's latent_preview (imported at module scope by the reference)."""


def get_previewer(*args, **kwargs):
    return None
