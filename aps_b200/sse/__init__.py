"""B200-native stand-ins for the `aps.sse` modules on the hot path."""
