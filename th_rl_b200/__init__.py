"""th_rl_b200 — B200-native implementation of th_rl's training hot path."""
__version__ = "0.1.0"
