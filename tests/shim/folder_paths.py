"""Stand-in for ComfyUI's folder_paths (imported at module scope by the reference)."""


def get_temp_directory():
    return "/tmp"


def get_save_image_path(prefix, output_dir, *args):
    return output_dir, prefix, 0, "", prefix
