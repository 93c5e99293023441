"""Reference package name (arch/) -> B200 implementations."""
