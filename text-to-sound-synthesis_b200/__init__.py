"""diffsound_b200 -- B200-native (sm_100a) kernels for the Diffsound text-to-sound hot path.

The directory is named ``text-to-sound-synthesis_b200`` (not importable by that name); load it through ``_pkg.load()`` at
the repository root, which registers it as the module ``diffsound_b200``.
"""
__version__ = "0.1.0"
