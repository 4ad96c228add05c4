"""B200-native batch engine for the CRI ADX / HCA codec hot paths."""
