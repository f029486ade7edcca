"""mkg_analogy_b200 — B200-native MarT (MKGformer) training hot path behind a C ABI."""
__version__ = "0.1.0"
