"""DDIMSampler with a fused classifier-free-guidance + DDIM update kernel.

Mirror of reference ldm/models/diffusion/ddim.py:14-243: `DDIMSampler(model, schedule="linear")`,
`.make_schedule(...)`, `.sample(S, batch_size, shape, conditioning, ..., eta, x_T,
unconditional_guidance_scale, unconditional_conditioning)` -> `(samples, {'x_inter', 'pred_x0'})`.
`model` must expose `.num_timesteps .device .betas .alphas_cumprod .alphas_cumprod_prev .apply_model(x, t, c)`
exactly as the reference requires; the alpha / sigma tables are read from it the way ddim.py:40-66 does
(so an fp16-rounded schedule of a `.half()` model is inherited).  Quirks kept: the uniform
schedule has S+1 or S+2 entries and `t_start=-1` drops the last one (76 / 49 iterations for S = 75 / 50).
Noise comes from torch's global generator in the reference's order: x_T first, then one randn per step.  One addition to
the reference's signature: `step_noise` (a sequence of tensors, one per iteration) replaces those per-step draws, which is
how several sampler calls of the reference run as ONE batched call here with the numbers each of them would have drawn
(zero123.generate_views).
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops_a as A


def make_ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps):
    c = num_ddpm_timesteps // num_ddim_timesteps
    return np.asarray(list(range(0, num_ddpm_timesteps, c))) + 1


class DDIMSampler(object):
    def __init__(self, model, schedule="linear", **kwargs):
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        self.device = model.device

    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        if ddim_discretize != "uniform":
            raise NotImplementedError("only the uniform discretisation is used by Zero123")
        self.ddim_timesteps = make_ddim_timesteps(ddim_num_steps, self.ddpm_num_timesteps)
        ac = self.model.alphas_cumprod.detach().to(torch.float32).cpu()
        assert ac.shape[0] == self.ddpm_num_timesteps
        self.alphas_cumprod = ac
        a = ac[self.ddim_timesteps].numpy().astype(np.float64)
        a_prev = np.asarray([float(ac[0])] + ac[self.ddim_timesteps[:-1]].tolist())
        sig = ddim_eta * np.sqrt((1 - a_prev) / (1 - a) * (1 - a / a_prev))
        self.ddim_alphas = torch.from_numpy(a.astype(np.float32))
        self.ddim_alphas_prev = torch.from_numpy(a_prev.astype(np.float32))
        self.ddim_sigmas = torch.from_numpy(sig.astype(np.float32))
        self.ddim_sqrt_one_minus_alphas = torch.from_numpy(np.sqrt(1. - a).astype(np.float32))

    @torch.no_grad()
    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0., mask=None, x0=None, temperature=1., noise_dropout=0., score_corrector=None,
               corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100, unconditional_guidance_scale=1.,
               unconditional_conditioning=None, dynamic_threshold=None, step_noise=None, **kwargs):
        if mask is not None or quantize_x0 or score_corrector is not None or noise_dropout > 0 or dynamic_threshold is not None:
            raise NotImplementedError("inpainting masks / quantisation / score correctors are not used by Zero123")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=verbose)
        C, H, W = shape
        return self.ddim_sampling(conditioning, (batch_size, C, H, W), x_T=x_T, callback=callback, img_callback=img_callback,
                                  log_every_t=log_every_t, temperature=temperature,
                                  unconditional_guidance_scale=unconditional_guidance_scale,
                                  unconditional_conditioning=unconditional_conditioning, step_noise=step_noise)

    @torch.no_grad()
    def ddim_sampling(self, cond, shape, x_T=None, callback=None, img_callback=None, log_every_t=100, temperature=1.,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, t_start=-1, step_noise=None, **kwargs):
        device = self.model.betas.device
        b = shape[0]
        img = torch.randn(shape, device=device) if x_T is None else x_T
        img = img.float().contiguous()
        timesteps = self.ddim_timesteps[:t_start]
        intermediates = {'x_inter': [img], 'pred_x0': [img]}
        total_steps = timesteps.shape[0]
        if step_noise is not None and len(step_noise) != total_steps:
            raise ValueError("step_noise holds %d tensors, the schedule has %d iterations" % (len(step_noise), total_steps))
        for i, step in enumerate(np.flip(timesteps)):
            index = total_steps - i - 1
            ts = torch.full((b,), int(step), device=device, dtype=torch.long)
            img, pred_x0 = self.p_sample_ddim(img, cond, ts, index=index, temperature=temperature,
                                              unconditional_guidance_scale=unconditional_guidance_scale,
                                              unconditional_conditioning=unconditional_conditioning,
                                              noise=None if step_noise is None else step_noise[i])
            if callback:
                img = callback(i, img, pred_x0)
            if img_callback:
                img_callback(pred_x0, i)
            if index % log_every_t == 0 or index == total_steps - 1:
                intermediates['x_inter'].append(img)
                intermediates['pred_x0'].append(pred_x0)
        return img, intermediates

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, temperature=1., unconditional_guidance_scale=1.,
                      unconditional_conditioning=None, noise=None, **kwargs):
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            e = self.model.apply_model(x, t, c)
            e2 = torch.cat([e, e]).float().contiguous()          # scale * (e - e) = 0: same formula, one kernel
            scale = 1.0
        else:
            x_in, t_in = torch.cat([x] * 2), torch.cat([t] * 2)
            # the conditioning does not change between the steps of one sampling call: build [uncond | cond] once and
            # hand the model the SAME tensor objects every step (lets it reuse context-only work, e.g. the UNet's
            # single-token cross-attention)
            cache = getattr(self, "_cin", None)
            if cache is not None and cache[0] is c and cache[1] is unconditional_conditioning:
                c_in = cache[2]
            else:
                if isinstance(c, dict):
                    c_in = {k: ([torch.cat([unconditional_conditioning[k][i], c[k][i]]) for i in range(len(c[k]))]
                                if isinstance(c[k], list) else torch.cat([unconditional_conditioning[k], c[k]])) for k in c}
                else:
                    c_in = torch.cat([unconditional_conditioning, c])
                self._cin = (c, unconditional_conditioning, c_in)
            e2 = self.model.apply_model(x_in, t_in, c_in).float().contiguous()
            scale = unconditional_guidance_scale
        if noise is None:
            noise = torch.randn(x.shape, device=x.device)                # noise_like(), reference util.py:264-267
        elif tuple(noise.shape) != tuple(x.shape):
            raise ValueError("step noise of shape %s for a state of shape %s" % (tuple(noise.shape), tuple(x.shape)))
        noise = noise.float().contiguous()
        if temperature != 1.:
            noise = noise * temperature
        return A.cfg_ddim_update(x.float().contiguous(), e2, noise, scale, float(self.ddim_alphas[index]),
                                 float(self.ddim_alphas_prev[index]), float(self.ddim_sigmas[index]),
                                 float(self.ddim_sqrt_one_minus_alphas[index]))
