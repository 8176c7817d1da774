"""Modules with the names and call signatures of the reference's compiled extensions."""
