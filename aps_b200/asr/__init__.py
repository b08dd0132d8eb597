"""B200-native stand-ins for the `aps.asr` modules on the hot path."""
