"""No-network stand-in for `ogb`: serves the seeded synthetic ARXIV-shape dataset."""
