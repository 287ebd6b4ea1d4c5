"""xspect2_b200 — B200-native k-mer scoring path of XspecT (COBS / Bloom), behind the reference's model API."""

__version__ = "0.1.0"
