"""Executes the source text of the reference's Rust host helpers on the LBM path (no Rust toolchain in the image)."""
