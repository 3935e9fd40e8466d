"""Factory + seeding — mirror of MuseDiffusion/utils/initialization.py:11-26,108-136 (the seam where the drop-in
installs).  Weight-overloading helpers of the reference (:29-87) are training-side and out of scope."""
import random
from types import SimpleNamespace

import numpy as np
import torch

from .diffusion import SpacedDiffusion, get_named_beta_schedule, space_timesteps
from .network import TransformerNetModel


def seed_all(seed, deterministic=False):
    """initialization.py:11-26: python / numpy / torch RNGs from hash(seed); the in-kernel Philox stream of this
    package takes its seed from torch.initial_seed() at loop entry, so it is covered as well."""
    if isinstance(seed, int):
        seed = hash(seed)
    random.seed(seed)
    from .corruption import generator                       # initialization.py:15,26: the corruptions' own stream
    generator.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def create_model_and_diffusion(args=None, **kwargs):
    """initialization.py:108-136: `create_model_and_diffusion(args) -> (model, diffusion)`, reading
    hidden_dim, hidden_t_dim, vocab_size, seq_len, dropout, noise_schedule, diffusion_steps, timestep_respacing,
    rescale_timesteps, predict_xstart from `args` (any attribute object; keywords also accepted).  The optional
    attribute `encoder_config` (dict) overrides the bert-base encoder sizes for the scaled-up benchmark config."""
    if args is None:
        args = SimpleNamespace(**kwargs)
    model = TransformerNetModel(input_dims=args.hidden_dim, output_dims=args.hidden_dim,
                                hidden_t_dim=args.hidden_t_dim, vocab_size=args.vocab_size, seq_len=args.seq_len,
                                dropout=args.dropout, encoder_config=getattr(args, "encoder_config", None))
    betas = get_named_beta_schedule(args.noise_schedule, args.diffusion_steps)
    timestep_respacing = getattr(args, "timestep_respacing", "")
    if not timestep_respacing:
        timestep_respacing = [args.diffusion_steps]
    diffusion = SpacedDiffusion(use_timesteps=space_timesteps(args.diffusion_steps, timestep_respacing), betas=betas,
                                rescale_timesteps=args.rescale_timesteps, predict_xstart=args.predict_xstart)
    return model, diffusion
