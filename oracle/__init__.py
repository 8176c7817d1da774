"""ORACLE package — test infrastructure only (see oracle/README.md). Not importable from the product path."""
