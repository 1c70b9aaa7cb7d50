"""joint-tensorf_b200: B200-native (sm_100a) TensoRF-VM volume-rendering hot path."""
