"""Stand-in for ComfyUI custom_nodes namespace (the reference probes it for optional integrations)."""
