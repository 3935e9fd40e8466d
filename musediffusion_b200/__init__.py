"""musediffusion_b200 — B200-native MuseDiffusion reverse-diffusion sampling path."""
