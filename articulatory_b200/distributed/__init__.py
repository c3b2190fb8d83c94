"""One process per GPU: launcher + data-parallel plumbing (replaces reference articulatory/distributed)."""
