"""Mutation check of the golden tests made from the reference's source: each entry breaks ONE
line of the host logic (argument order, a slice, a weight, a key, the shard order ...), runs the
test that is supposed to pin it and restores the file.  Every mutation must be DETECTED.

    python tools/mutation_check.py        (CPU only; a few minutes)
"""
import os
os.chdir(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import subprocess, sys
def mutate(path, old, new, test):
    s=open(path).read()
    assert old in s, (path, old)
    open(path,"w").write(s.replace(old,new,1))
    try:
        r=subprocess.run([sys.executable,"-m","pytest",*test.split(),"-q","-x","-m","not gpu"],capture_output=True,text=True)
        last=r.stdout.strip().splitlines()[-1]
    finally:
        open(path,"w").write(s)
    print(("DETECTED " if "failed" in last or "error" in last else "MISSED   ")+f"{path}: {old[:50]!r} -> {new[:50]!r}   [{last}]")
B="sup3r_b200/models/base.py"; A="sup3r_b200/models/abstract.py"
mutate(B,"return self.loss_fun(hi_res_gen, hi_res_true)","return self.loss_fun(hi_res_true, hi_res_gen)","tests/test_gan_loss_golden.py")
mutate(B,"loss_gen_advers = self.calc_loss_disc(disc_out_gen, disc_out_true)","loss_gen_advers = self.calc_loss_disc(disc_out_true, disc_out_gen)","tests/test_gan_loss_golden.py")
mutate(B,"crop = [(0, 0)] * (hi_res_gen.dim() - 1) + [(0, n_exo)]","crop = [(0, 0)] * (hi_res_gen.dim() - 1) + [(0, 0)]","tests/test_gan_loss_golden.py")
mutate(A,"term = val if w == 1.0 else val * w","term = val","tests/test_gan_loss_golden.py")
mutate(A,'nm = exo_name.replace("_obs", "") if exo_name not in self._means else exo_name','nm = exo_name',"tests/test_gan_loss_golden.py")
mutate(B,"if compute_disc or train_disc:","if train_disc:","tests/test_gan_loss_golden.py")
mutate("sup3r_b200/parallel.py","src = world_size() - 1 if src is None else src","src = 0 if src is None else src","tests/test_parallel_golden.py")
mutate("sup3r_b200/parallel.py","return x[index * k:(index + 1) * k]","return x[(n_shards - 1 - index) * k:(n_shards - index) * k]","tests/test_parallel_golden.py")
mutate(B,"epochs = [e + int(self._history.index.values[-1]) + 1 for e in epochs]","epochs = [e + int(self._history.index.values[-1]) for e in epochs]","tests/test_train_loop_golden.py")
mutate(B,'extras.update({f"OptmDisc/{k}": v for k, v in opt_d.items()})','extras.update({f"OptmDisc/{k}": v for k, v in opt_g.items()})',"tests/test_train_loop_golden.py")
mutate("sup3r_b200/loss_metrics.py","return t[:, :, :, ::self._t_enhance, :]","return t[:, :, :, 1::self._t_enhance, :]","tests/test_losses_golden.py")
mutate("sup3r_b200/loss_metrics.py","mmd = mmd - torch.mean(2 * gaussian_kernel(x1, x2, sigma))","mmd = mmd - torch.mean(gaussian_kernel(x1, x2, sigma))","tests/test_losses_golden.py")
# ---- pins made earlier in the round
mutate("sup3r_b200/models/dc.py",'"spatial_weights": share(total.mean(axis=1))','"spatial_weights": share(total.mean(axis=1) ** 2)',"tests/test_training_schedule_golden.py")
mutate("sup3r_b200/models/dc.py","*divmod(i, n_t))]","*divmod(i + 1, n_t))]","tests/test_training_schedule_golden.py")
mutate("sup3r_b200/models/multi_step.py","np.transpose(hi_res, axes=(1, 2, 0, 3))[np.newaxis]","np.transpose(hi_res, axes=(2, 1, 0, 3))[np.newaxis]","tests/test_multistep_golden.py")
mutate("sup3r_b200/models/solar_cc.py","p0 = (24 - self.POINT_LOSS_HOURS) // 2","p0 = (24 - self.POINT_LOSS_HOURS) // 2 + 1","tests/test_solar_golden.py")
mutate("sup3r_b200/models/solar_cc.py","term = (c_sub + c_24h) / nd","term = (c_sub + c_24h)","tests/test_solar_golden.py")
mutate("sup3r_b200/pipeline/strategy.py","t_enhance = int(n_hr / len(lr))","t_enhance = int(n_hr / len(lr)) + 1","tests/test_grid_golden.py")
mutate("sup3r_b200/pipeline/strategy.py","lon[:, -1], lat[:, -1] = lon[:, -2] + d_right, lat[:, -2]","lon[:, -1], lat[:, -1] = lon[:, -2], lat[:, -2]","tests/test_grid_golden.py")
mutate("sup3r_b200/pipeline/strategy.py","n = int(min(self.max_nodes or np.inf, max(len(chunks), 1)))","n = int(min((self.max_nodes or np.inf) + 1, max(len(chunks), 1)))","tests/test_strategy_golden.py")
mutate("sup3r_b200/pipeline/strategy.py","start = 0 if not padded.start else self.temporal_pad","start = 0","tests/test_strategy_golden.py")
mutate("sup3r_b200/models/with_obs.py","obs_frac = float(np.float32(int((~mask).sum())) / np.float32(mask.numel()))","obs_frac = float(np.float32(int((mask).sum())) / np.float32(mask.numel()))","tests/test_obs_golden.py")
mutate("sup3r_b200/models/abstract.py","def early_stop(history, column, threshold=0.005, n_epoch=5):","def early_stop(history, column, threshold=0.005, n_epoch=5):\n        n_epoch += 1","tests/test_training_schedule_golden.py")
# ---- slicer, forward-pass helpers, coarsening, bias, exo data, interface, post-processing
mutate("sup3r_b200/pipeline/slicer.py","stop = None if self.spatial_pad == 0 else -start","stop = None","tests/test_host_golden.py")
mutate("sup3r_b200/pipeline/forward_pass.py","            n = model_step + 1","            n = model_step","tests/test_forward_pass_golden.py")
mutate("sup3r_b200/utilities.py",'"total": np.nansum, "max": np.max','"total": np.nansum, "max": np.min',"tests/test_host_golden.py")
mutate("sup3r_b200/utilities.py","return data[:, :, :, ::t_enhance, :]","return data[:, :, :, 1::t_enhance, :]","tests/test_host_golden.py")
mutate("sup3r_b200/bias.py","scalar, adder = _smooth(scalar, smoothing), _smooth(adder, smoothing)\n    return _clip(data * scalar + adder, out_range)","scalar, adder = _smooth(scalar, smoothing), adder\n    return _clip(data * scalar + adder, out_range)","tests/test_postprocess.py")
mutate("sup3r_b200/exo.py",'new[k] = v[tuple(sl)[: len(v.shape) - 1]] if k == "data" else v','new[k] = v[tuple(sl)[: len(v.shape) - 2]] if k == "data" else v',"tests/test_host_golden.py tests/test_multistep_golden.py")
mutate("sup3r_b200/exo.py",'if min_step <= s["model"] and (max_step is None or s["model"] < max_step)]','if min_step <= s["model"] and (max_step is None or s["model"] <= max_step)]',"tests/test_host_golden.py tests/test_multistep_golden.py")
mutate("sup3r_b200/models/interface.py",'obs = [f.replace("_obs", "") for f in self.obs_features]','obs = [f for f in self.obs_features]',"tests/test_interface_golden.py")
mutate("sup3r_b200/models/abstract.py","hi_res_exo = np.repeat(np.expand_dims(hi_res_exo, 3), hi_res.shape[3], axis=3)","hi_res_exo = np.repeat(np.expand_dims(hi_res_exo, 3), hi_res.shape[3] + 1, axis=3)","tests/test_gan_loss_golden.py")
mutate("sup3r_b200/pipeline/postprocess.py","dx = (dx + 180) % 360 - 180","dx = (dx + 180) % 360","tests/test_postprocess.py")
# ---- training schedule boundaries, history bookkeeping, output check
A="sup3r_b200/models/abstract.py"; B="sup3r_b200/models/base.py"; F="sup3r_b200/pipeline/forward_pass.py"; S="sup3r_b200/pipeline/strategy.py"
T="tests/test_training_schedule_golden.py"
mutate(A,"return record.iloc[-max_batches:]","return record.iloc[-(max_batches + 1):]",T)
mutate(A,"key = k if prefix is None or prefix in k else prefix + k","key = k if prefix is None else prefix + k",T)
mutate(A,"chp = checkpoint_int is not None and (epoch % checkpoint_int) == 0","chp = checkpoint_int is not None and (epoch % checkpoint_int) == 1",T)
mutate(B,"disc_too_good = loss_disc <= disc_th_low","disc_too_good = loss_disc < disc_th_low",T)
# (dropping "and train_disc" from disc_too_bad is an equivalent mutant: without discriminator
#  training the generator step is unconditional, base.py:1008-1031)
mutate(B,"if only_gen or (train_gen and not gen_too_good):","if only_gen or train_gen:",T)
mutate(F,"if chk[i, 0] == chk[i, 1] and chk[i, 0] not in allowed_const:","if chk[i, 0] == chk[i, 1]:","tests/test_host_golden.py tests/test_forward_pass_golden.py")
mutate(F,"if allowed_const is True:\n            return False\n        if allowed_const is False or allowed_const is None:","if allowed_const is True:\n            return True\n        if allowed_const is False or allowed_const is None:","tests/test_host_golden.py tests/test_forward_pass_golden.py")
mutate(A,"nm = ","nm = exo_name  # ","tests/test_gan_loss_golden.py")
# ---- incremental restart, exo padding, feature combination, resolution / enhancement checks,
# ---- collector stitching, drop_leap, _sum_parallel_grad
F="sup3r_b200/pipeline/forward_pass.py"; S="sup3r_b200/pipeline/strategy.py"; I="sup3r_b200/models/interface.py"; M="sup3r_b200/models/multi_step.py"
mutate(S,"done = out_file is not None and os.path.exists(out_file) and self.incremental","done = out_file is not None and os.path.exists(out_file)","tests/test_strategy_golden.py")
mutate(F,"for en, pw in zip([s_en, s_en, t_en], pad_width)), (0, 0))","for en, pw in zip([s_en, s_en, 1], pad_width)), (0, 0))","tests/test_forward_pass_golden.py")
mutate(F,'step["t_enhance"] * input_data.shape[2], axis=2)','input_data.shape[2], axis=2)',"tests/test_forward_pass_golden.py")
mutate(I,"exo_feats = [] if n_missing <= 0 else features[-n_missing:]","exo_feats = [] if n_missing <= 0 else features[:n_missing]","tests/test_interface_golden.py tests/test_multistep_golden.py")
mutate(I,'ok = ires["temporal"] / ores["temporal"] == t and ires["spatial"] / ores["spatial"] == s','ok = ires["temporal"] / ores["temporal"] == t or ires["spatial"] / ores["spatial"] == s',"tests/test_interface_golden.py tests/test_reference_host_cases.py")
mutate(I,"ls = ls if ls is not None else s","ls = s","tests/test_interface_golden.py tests/test_reference_host_cases.py")
mutate("sup3r_b200/pipeline/writers.py","rows, cols = gids // full_shape[1], gids % full_shape[1]","rows, cols = gids // full_shape[0], gids % full_shape[0]","tests/test_writers.py")
mutate("sup3r_b200/bias.py","if drop_leap:","if not drop_leap:","tests/test_postprocess.py")
mutate("sup3r_b200/models/abstract.py","grad, loss_details = future.result()","grad, _ = future.result()","tests/test_parallel_golden.py")
# ---- multi-step flag / exo routing, observation masks, the oracles themselves
M="sup3r_b200/models/multi_step.py"; W="sup3r_b200/models/with_obs.py"; O="oracle/losses_ref.py"
mutate(M,"i_norm_in = not (i == 0 and not norm_in)","i_norm_in = norm_in","tests/test_multistep_golden.py")
mutate(M,"i_un_norm_out = not (last and not un_norm_out)","i_un_norm_out = un_norm_out","tests/test_multistep_golden.py")
mutate(M,"exogenous_data.get_model_step_exo(i)","exogenous_data.get_model_step_exo(0)","tests/test_multistep_golden.py")
mutate(M,"hi_res = hi_res[..., [out_feats.index(fn) for fn in in_feats]]","hi_res = hi_res[..., :len(in_feats)]","tests/test_multistep_golden.py")
mutate(W,"t_mask = RANDOM_GENERATOR.uniform(size=mask_shape[-2]) <= time_frac","t_mask = RANDOM_GENERATOR.uniform(size=mask_shape[-2]) < time_frac - 0.3","tests/test_obs_golden.py")
mutate(W,"obs_mask = np.where(np.asarray(topo)[..., None] > 0, obs_mask, offshore_mask)","obs_mask = np.where(np.asarray(topo)[..., None] < 0, obs_mask, offshore_mask)","tests/test_obs_golden.py")
mutate(O,"return t.sum(axis=(2, 4)) / s ** 2","return t.sum(axis=(2, 4)) / s","tests/test_losses_golden.py")
mutate("oracle/torch_ref.py","logits = torch.cat([out_true - out_gen.mean(), out_gen - out_true.mean()], dim=0)","logits = torch.cat([out_true - out_true.mean(), out_gen - out_gen.mean()], dim=0)","tests/test_gan_loss_golden.py")
