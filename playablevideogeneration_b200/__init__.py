"""B200-native CADDY hot path (PlayableVideoGeneration): hand-written sm_100a CUDA behind the reference's Python
module protocol.  See DESIGN.md.  Importing this package never touches the GPU; the first op call loads
``libpvg_b200.so`` and fails loudly if it is missing (there is NO CPU or eager-PyTorch fallback)."""

__version__ = "0.1.0"
