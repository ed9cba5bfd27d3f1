"""WGSL-subset transpiler + runtime used to execute the reference's own shader source (test infrastructure)."""
