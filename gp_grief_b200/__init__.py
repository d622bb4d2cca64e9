"""B200-native GP-GRIEF hot path behind the gp_grief Python API."""
