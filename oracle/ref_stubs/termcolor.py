"""Test-infrastructure stub (oracle only): lets the read-only reference import here.
Console colouring is not on the arithmetic path."""


def colored(text, *args, **kwargs):
    return text
