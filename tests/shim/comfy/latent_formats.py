class LatentFormat:
    scale_factor = 1.0
    latent_channels = 4


class SD15(LatentFormat):
    scale_factor = 0.18215
