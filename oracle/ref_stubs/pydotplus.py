"""Test-infrastructure stub (oracle only)."""
