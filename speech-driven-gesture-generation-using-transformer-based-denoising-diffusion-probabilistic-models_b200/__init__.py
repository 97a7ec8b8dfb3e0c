"""B200-native DDPM gesture sampling: drop-in `create_model` / `Generator` over sm_100a kernels."""
__all__ = ["_lib"]
