"""BLIP2_MR (Mr. BLIP / Chrono) behind the LAVIS model surface, computed by the sm_100a kernels.

Drop-in for lavis/models/blip2_mr_models/blip2_mr.py: registry name "blip2_mr", from_config keys
(:1420-1464), forward(samples) -> {"loss"} (:300-307, :433-570), generate(samples, ...) -> dict
(:826-946), attribute / state-dict names (visual_encoder.*, ln_vision.*, Qformer.bert.*,
query_tokens, t5_proj.*, t5_model.base_model.model.* with peft-style lora_A/lora_B.default keys).

forward() in training mode runs forward AND backward kernels in one sweep and returns a loss whose
autograd node hands the finished gradients to .backward() (so GradScaler, DDP reducer hooks and
gradient accumulation in lavis/tasks/base_task.py:66-68,157-248 work unchanged).
There is no CPU or eager-PyTorch fallback: without a CUDA device / built extension, forward raises.
"""
import logging
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops, mr_utils
from .base_model import BaseModel, attach, disabled_train
from . import qa
from .dims import ANSWERER_PREFIX, Dims, FULL, T5_PREFIX, add_answerer, init_state_dict
from .registry import registry
from .t5 import T5Engine
from .tokenizer import load_t5_tokenizer
from .vision import VitEngine, QFormerEngine

INT_MIN = -2 ** 31


class _HandOverGrads(torch.autograd.Function):
    """loss = f(params) where df/dparams was already computed by the backward kernels into ONE flat fp32 buffer
    (BLIP2_MR._gflat).  backward() scales it by the incoming gradient (GradScaler / accumulation factor) with one kernel
    and returns per-parameter views, so autograd / DDP hooks / the optimiser see ordinary .grad tensors."""

    @staticmethod
    def forward(ctx, loss, model, *params):
        ctx.model = model
        return loss.clone()

    @staticmethod
    def backward(ctx, g):
        return (None, None) + tuple(ctx.model._hand_over(g))


class _GraphedStep:
    """Static input buffers + the captured CUDA graph of the device half of one training step, for one shape signature
    (clips, frames, encoder length, decoder length).  All small integer inputs travel in ONE pinned int32 buffer."""

    def __init__(self, b, t, Le, Ld, img_size, video_dtype=torch.float32, front=None):
        n = 2 * b * Le + 3 * b * Ld
        self.host = torch.empty(n, dtype=torch.int32).pin_memory()
        self.dev = torch.empty(n, dtype=torch.int32, device="cuda")
        self.front = front                                   # _GraphedFront: the text-independent front runs as its own graph
        self.video = front.video if front is not None else \
            torch.empty((b, t, 3, img_size, img_size), dtype=video_dtype, device="cuda")
        self.loss = torch.zeros((1,), dtype=torch.float32, device="cuda")
        self.h, self.d = {}, {}
        off = 0
        for name, cnt, shape in (("idx", b * Le, (b * Le,)), ("kmask", b * Le, (b, Le)), ("labels", b * Ld, (b, Ld)),
                                 ("dec_ids", b * Ld, (b, Ld)), ("dmask", b * Ld, (b, Ld))):
            self.h[name] = self.host[off:off + cnt].view(shape)
            self.d[name] = self.dev[off:off + cnt].view(shape)
            off += cnt
        self.graph = None
        self.calls = 0
        self.n_launch = 0
        self.staged = None                                   # event: the previous step's H2D copy of the pinned buffer has run

    def stage(self, host, video):
        if self.staged is not None:
            self.staged.synchronize()                        # the pinned buffer is still the source of an in-flight async copy
        for name in self.h:
            self.h[name].copy_(torch.from_numpy(host[name]).view(self.h[name].shape))
        self.dev.copy_(self.host, non_blocking=True)
        self.staged = torch.cuda.Event()
        self.staged.record()
        if video is not None:                                # None: the front graph's own buffer already holds the frames
            self.video.copy_(video.reshape(self.video.shape), non_blocking=True)


class _GraphedFront:
    """The part of a training step that does not depend on the text -- t5_proj re-cast, ViT, ln_vision, Q-Former, t5_proj
    (-> mean) -- as its OWN captured graph with its own static frame buffer, for one (clips, frames) signature.  It is launched
    as soon as the frames are on their way to the device, BEFORE the host builds the prompt table (tokeniser, interleave rows:
    ~3 ms of Python), so that work runs under the ViT instead of in front of the step -- which is where it sits whenever the
    caller reads the loss every iteration, as the reference's loop does (base_task.py:_train_inner_loop, loss.item())."""

    def __init__(self, b, t, img_size, video_dtype=torch.float32):
        self.video = torch.empty((b, t, 3, img_size, img_size), dtype=video_dtype, device="cuda")
        self.out = None                                      # (frames_for_t5, Q-Former hidden fp32, fp16): static once captured
        self.graph = None
        self.n_launch = 0


class Blip2Base(BaseModel):
    """blip2.py:27-119 surface (builders are replaced by the seeded / checkpoint-loaded parameter tree)."""


@registry.register_model("blip2_mr")
class BLIP2_MR(Blip2Base):
    PRETRAINED_MODEL_CONFIG_DICT = {
        "pretrain_flant5xl": "configs/models/blip2/blip2_pretrain_flant5xl.yaml",
    }

    def __init__(self, img_size=224, drop_path_rate=0, use_grad_checkpoint=False, vit_precision="fp16",
                 freeze_vit=True, num_query_token=32, t5_model="google/flan-t5-xl", num_beams=5, prompt="",
                 max_txt_len=200, apply_lemmatizer=False, input_time_format="seconds_integers",
                 interleave_data=True, frame_token_aggregation=None, task="qformer_freeze_lora",
                 num_frames_for_answer=4, resample_frames=False, dims: Dims = None, init_seed=1234,
                 lora_b_std=0.0, state_dict=None, tokenizer=None, cuda_graphs=True, graph_bucket=(32, 8),
                 train_dropout=None, dropout_seed=0, allow_synthetic=True):
        super().__init__()
        self.dims = d = dims or FULL
        assert img_size == d.img_size and num_query_token == d.num_query
        # two-stage video QA (blip2_mr.py:104-109): a frozen localizer (t5_model) proposes a moment, a second LoRA T5 answers
        self.use_localizer = "with_localizer" in task
        self.use_oracle_localizer = "oracle_localizer" in task
        self.num_frames_for_answer = num_frames_for_answer
        self.resample_frames = resample_frames
        self.video_processor_answerer_eval = None
        if "QA" in task and resample_frames:                 # blip2_mr.py:111-116: the answerer's own eval processor (n frames)
            from .data import VideoProcessor
            self.video_processor_answerer_eval = VideoProcessor(image_size=img_size, n_frms=num_frames_for_answer)
        if "lora" not in task:
            raise NotImplementedError("only LoRA tasks are implemented (every mr_BLIP yaml uses qformer_freeze_lora)")
        if not freeze_vit or "qformer_freeze" not in task:
            # the reference would train the ViT / the Q-Former here; this path has forward kernels only for both (every mr_BLIP
            # recipe freezes them: SURVEY.md section 3.1) -- refuse instead of silently training something else
            raise NotImplementedError("freeze_vit=False / a task without 'qformer_freeze' needs ViT / Q-Former backward kernels, "
                                      "which this path does not have")
        # use_grad_checkpoint (eva_vit.py:352: torch.utils.checkpoint over the ViT blocks) only trades memory for time in a ViT
        # that stores activations; the frozen ViT here stores none, so the flag is accepted and changes nothing
        self.use_grad_checkpoint = bool(use_grad_checkpoint)
        self.task = task
        self.use_lora = True
        self.post_process = mr_utils.post_process
        self.input_time_format = input_time_format
        self.interleave_data = interleave_data
        self.frame_token_aggregation = frame_token_aggregation
        self.num_query_token = num_query_token
        self.max_txt_len = max_txt_len
        self.num_beams = num_beams
        self.vit_precision = vit_precision

        # ---- parameters under the reference's names ------------------------------------------------
        sd = state_dict if state_dict is not None else init_state_dict(d, seed=init_seed, lora_b_std=lora_b_std)
        qa_task = "QA" in task
        if qa_task and not any(k.startswith(ANSWERER_PREFIX) for k in sd):
            sd = add_answerer(dict(sd), d, seed=init_seed, lora_b_std=lora_b_std, device=next(iter(sd.values())).device)
        shared, shared_ans = None, None
        for name, t in sd.items():
            if name == "query_tokens":
                attach(self, name, t.clone(), requires_grad="qformer_freeze" not in task)
                continue
            if name.startswith((T5_PREFIX, ANSWERER_PREFIX)) and name.endswith("embed_tokens.weight"):
                continue                                     # tied to shared.weight below
            if name.startswith(ANSWERER_PREFIX) and not qa_task:
                continue
            lora = "lora_A" in name or "lora_B" in name
            if qa_task:                                      # the localizer's adapters are frozen, the answerer's train (blip2_mr.py:199-235)
                lora = lora and name.startswith(ANSWERER_PREFIX)
            train = (lora or name.startswith("t5_proj.")
                     or (name.startswith("Qformer.") and "qformer_freeze" not in task))
            t = t.clone()
            if name.startswith("visual_encoder.") and vit_precision == "fp16" and t.ndim >= 2 and "pos_embed" not in name \
                    and "cls_token" not in name:
                t = t.half()                                 # convert_weights_to_fp16, eva_vit.py:397-412
            if name.startswith("visual_encoder.") and name.endswith((".fc1.bias", ".fc2.bias", ".proj.bias")) \
                    and vit_precision == "fp16":
                t = t.half()
            p = attach(self, name, t, requires_grad=train and t.is_floating_point())
            if name == T5_PREFIX + "shared.weight":
                shared = p
            if name == ANSWERER_PREFIX + "shared.weight":
                shared_ans = p
        for side in ("encoder", "decoder"):                  # HF ties embed_tokens to shared
            attach(self, T5_PREFIX + side + ".embed_tokens.weight", shared)
            if shared_ans is not None:
                attach(self, ANSWERER_PREFIX + side + ".embed_tokens.weight", shared_ans)
        if freeze_vit:
            self.visual_encoder.eval()
            self.visual_encoder.train = disabled_train.__get__(self.visual_encoder)
            logging.info("freeze vision encoder")

        # ---- tokenizer + number-token hygiene (blip2_mr.py:143,165-168) ----------------------------
        self.t5_tokenizer = tokenizer or load_t5_tokenizer(t5_model, allow_synthetic=allow_synthetic)
        self.annoying_numbers, _ = mr_utils.find_annoying_numbers(self.t5_tokenizer, 200)
        self.annoying_numbers_replacement_dict = mr_utils.find_annoying_numbers_replacement_dict(self.annoying_numbers)
        self.seperator_token = self.t5_tokenizer.convert_tokens_to_ids(">")
        self.pad_token_id = self.t5_tokenizer.pad_token_id
        self._engines = None
        self._answerer = None                                # QA branch: T5Engine over answerer_model.*
        self._lora_versions = None
        # Training steps replay a captured CUDA graph per shape signature (first sight of a shape runs eagerly, the second
        # captures).  Encoder / decoder lengths are padded (masked, exact) up to graph_bucket so few graphs cover a dataset.
        self.cuda_graphs = cuda_graphs and os.environ.get("MRB_CUDA_GRAPHS", "1") != "0"
        self.graph_bucket = graph_bucket
        self.max_graphs = 16
        self._steps = {}
        self._seen = {}                                      # shape signature -> times seen (survives LRU eviction)
        self._graph_pool = None
        self._in_device_step = False
        self._defer_refresh = False                          # the captured step re-packs LoRA / t5_proj itself
        # MRB_SPLIT_GRAPH=1: front (ViT .. t5_proj) and back (T5) of the step as two graphs, the host's prompt assembly in between
        self.split_graph = os.environ.get("MRB_SPLIT_GRAPH", "0") == "1"
        self._fronts = {}
        # Train-mode dropout (the reference's train() step: Q-Former / T5 0.1, LoRA inputs 0.05; mr_blip_b200/dropout.py): ON by
        # default, as in the reference, where module.train() switches every nn.Dropout on.  train_dropout=False (or
        # MRB_TRAIN_DROPOUT=0) makes train() steps run the eval() arithmetic -- rate 0 -- which is what the parity tests of the
        # hand-written backward against the (mask-free) golden gradients use.
        self.train_dropout = (os.environ.get("MRB_TRAIN_DROPOUT", "1") != "0") if train_dropout is None else bool(train_dropout)
        self.dropout_seed = int(dropout_seed)
        self.drop_state = None

    # ---------------------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, cfg):
        """Same keys as blip2_mr.py:1420-1464.  The third-party pieces the reference fetches from hubs (tokenizer, FlanT5, EVA ViT,
        BLIP-2 / Mr. BLIP checkpoints) must be LOCAL files here (no network).  When one of them is missing this raises -- a run on the
        synthetic tokenizer and seeded random weights would train and evaluate on garbage without ever failing -- unless the recipe says
        `allow_synthetic: true` (dry runs, tests, benchmarks)."""
        get = cfg.get
        synth = bool(get("allow_synthetic", False))
        if not synth:
            t5 = get("t5_model", "google/flan-t5-xl")
            if not (isinstance(t5, str) and os.path.isdir(t5)):
                raise RuntimeError("model.t5_model=%r is not a local FlanT5 directory: the T5 weights would stay seeded random values. "
                                   "Set model.t5_model to a local directory, or model.allow_synthetic: true for a dry run" % (t5,))
        model = cls(img_size=get("image_size", 224), drop_path_rate=get("drop_path_rate", 0),
                    use_grad_checkpoint=get("use_grad_checkpoint", False), vit_precision=get("vit_precision", "fp16"),
                    freeze_vit=get("freeze_vit", True), num_query_token=get("num_query_token", 32),
                    t5_model=get("t5_model", "google/flan-t5-xl"), num_beams=get("num_beams", 5), prompt=get("prompt", ""),
                    max_txt_len=get("max_len", 200), apply_lemmatizer=get("apply_lemmatizer", False),
                    input_time_format=get("input_time_format", "seconds_integers"),
                    interleave_data=get("interleave_data", True), frame_token_aggregation=get("frame_token_aggregation", None),
                    task=get("task", "qformer_freeze_lora"), num_frames_for_answer=get("num_frames_for_answer", 4),
                    resample_frames=get("resample_frames", False), dims=get("dims", None),
                    init_seed=get("init_seed", 1234), lora_b_std=get("lora_b_std", 0.0),
                    train_dropout=get("train_dropout", None), dropout_seed=get("dropout_seed", 0), allow_synthetic=synth)
        # third-party weights the reference fetches from hubs, from local files here: `t5_model` may be a transformers
        # directory (as it may be for from_pretrained), `vit_weights` an eva_vit_g.pth
        from . import weights
        if isinstance(get("t5_model", None), str) and os.path.isdir(get("t5_model")):
            weights.load_hf_t5(model, get("t5_model"))
        if get("vit_weights", None):
            weights.load_eva_vit(model, get("vit_weights"))
        elif not synth:
            logging.error("model.vit_weights is not set: the EVA ViT-g keeps seeded random weights unless the `pretrained` checkpoint "
                          "carries visual_encoder.* (the BLIP-2 checkpoints do not)")
        model.load_checkpoint_from_config(cfg, allow_synthetic=synth)
        return model

    def load_checkpoint_from_config(self, cfg, allow_synthetic=True, **kwargs):
        """blip2_mr.py:1466-1495: pretrained BLIP-2 weights, then the finetuned LoRA adapter (local files only).  A configured
        checkpoint that is not a local file is an error unless allow_synthetic."""
        for key in ("pretrained", "finetuned") if cfg.get("load_finetuned", True) else ("pretrained",):
            path = cfg.get(key, None)
            if path and os.path.isfile(path):
                self.load_checkpoint(path)
            elif path and not allow_synthetic:
                raise RuntimeError("model.%s=%r is not a local file (no network here); set model.allow_synthetic: true to keep the "
                                   "seeded weights for a dry run" % (key, path))
            elif path:
                logging.warning("%s checkpoint %s is not a local file; keeping seeded weights", key, path)

    def reset_graphs(self):
        """Drop every captured step graph (and the private memory pool they share: a pool handle is only valid while
        one of its graphs is alive)."""
        self._steps, self._seen, self._graph_pool, self._fronts = {}, {}, None, {}

    def _weights_changed(self):
        self._engines = None
        self.reset_graphs()

    # ---------------------------------------------------------------------------------------------
    def _get(self, name):
        obj = self
        for p in name.split("."):
            obj = obj._modules[p] if p in obj._modules else obj._parameters[p]
        return obj

    def engines(self):
        if self._engines is None:
            if not torch.cuda.is_available() or self.device.type != "cuda":
                raise RuntimeError("BLIP2_MR runs only on a CUDA device through libmrblip_b200.so "
                                   "(no CPU / eager fallback); move the model with .cuda()")
            d = self.dims
            vit, qf, t5 = VitEngine(d, self._get), QFormerEngine(d, self._get), T5Engine(d, self._get)
            self._engines = (vit, qf, t5)
            qa_task = "QA" in self.task
            self._answerer = T5Engine(d, self._get, prefix=ANSWERER_PREFIX) if qa_task else None
            self._lora_versions = None
            self.reset_graphs()
            # every trainable gradient lives in one flat fp32 buffer: zeroed / scaled / all-reduced with single launches
            pw, pb = self.t5_proj.weight, self.t5_proj.bias
            tr = self._answerer if qa_task else t5           # the T5 whose adapters train
            n = tr.n_grad_elems()
            # QA: the frame embeddings are computed under no_grad (blip2_mr.py:316-372), so t5_proj gets no gradient there
            self._gflat = torch.zeros(n + (0 if qa_task else pw.numel() + pb.numel()), dtype=torch.float32, device="cuda")
            self._hflat = torch.empty_like(self._gflat)
            off = tr.bind_grads(self._gflat, 0)
            assert off == n
            if qa_task:
                self._grad_params = [p for p, _ in tr.param_grads()]
            else:
                self._g_projW = self._gflat[n:n + pw.numel()].view(pw.shape)
                self._g_projb = self._gflat[n + pw.numel():].view(pb.shape)
                self._grad_params = [p for p, _ in t5.param_grads()] + [pw, pb]
        vit, qf, t5 = self._engines
        if self.train_dropout and self.drop_state is None:
            from .dropout import DropState
            import torch.distributed as tdist
            rank = tdist.get_rank() if (tdist.is_available() and tdist.is_initialized()) else 0
            self.drop_state = DropState(base_seed=self.dropout_seed + rank)      # per-rank masks (reference: train.py:57-58, seed + get_rank())
        if self._in_device_step or self._defer_refresh:      # (captured) device step: LoRA re-pack is part of the step itself
            return vit, qf, t5
        tr = self._answerer if self._answerer is not None else t5
        vers = tuple(p._version for g in tr.groups for p in g.A_params + g.B_params)
        vers += (self.t5_proj.weight._version, self.t5_proj.bias._version)
        if vers != self._lora_versions:                      # optimizer.step() happened: re-pack the trainable bits
            tr.refresh()
            qf.set_t5_proj(self.t5_proj.weight, self.t5_proj.bias)
            self._lora_versions = vers
        return vit, qf, t5

    # ---------------------------------------------------------------------------------------------
    def get_frame_embeddings_and_attentions(self, image, want_aux=False):
        """blip2_mr.py:948-988: ViT -> ln_vision -> Q-Former -> t5_proj (-> mean).  image [b,t,3,H,W].
        -> frames_for_t5 fp32 [b, t*n, 2048], ones [b, t*n]."""
        vit, qf, _ = self.engines()
        if isinstance(image, list):
            image = torch.stack(image)
        b, t = image.shape[:2]
        img = image.reshape(b * t, *image.shape[2:])
        # raw uint8 frames stay uint8 (normalisation is fused into the patch extraction); anything else is the reference's fp32
        img = img.to(device="cuda", non_blocking=True) if img.dtype == torch.uint8 else \
            img.to(device="cuda", dtype=torch.float32, non_blocking=True)
        with ops.phase("vit"):
            x = vit.forward(img)
        with ops.phase("qformer"):
            h, h16 = qf.forward(x, b * t)
            f = qf.project(h16)
        n = self.dims.num_query
        if self.frame_token_aggregation:
            assert self.frame_token_aggregation in ["mean"], "Invalid aggregation method, please choose from ['mean']"
            agg = torch.empty((b * t, self.dims.d_model), dtype=torch.float32, device="cuda")
            ops.group_mean(f, agg, b * t, n, self.dims.d_model)
            f, n = agg, 1
        atts = torch.ones((b, t * n), dtype=torch.long, device="cuda")
        if want_aux:
            return f.view(b, t * n, -1), atts, (h, h16)
        return f.view(b, t * n, -1), atts

    def _timestamps(self, timestamps, durations):
        fmt, table = self.input_time_format, self.annoying_numbers_replacement_dict
        fns = {"seconds_integers": mr_utils.get_timestamps_as_seconds_integers,
               "seconds_floats": mr_utils.get_timestamps_as_seconds_floats,
               "relative_integers": mr_utils.get_timestamps_as_relative_integers,
               "relative_floats": mr_utils.get_timestamps_as_relative_floats,
               "framenumbers": mr_utils.get_timestamps_as_framenumbers}
        if fmt not in fns:
            raise ValueError("Invalid input_time_format, please choose from ['framenumbers', 'relative_floats', "
                             "'relative_integers', 'seconds_integers', 'seconds_floats']")      # blip2_mr.py:627-630
        return fns[fmt](timestamps, durations, table)

    def _clean_ids(self, values):
        """get_clean_timestamp_tokens_and_embs (blip2_mr.py:1561-1608): ids without specials, leading '▁' (3) dropped."""
        ids = self.t5_tokenizer([str(v) for v in values], add_special_tokens=False)["input_ids"]
        return [i[1:] if (len(i) > 1 and i[0] == 3) else i for i in ids]

    def build_prompt_table(self, timestamps, durations, B, T, n, video_prompt_end, query_prompt, task_prompt):
        """Host half of prompt_concatenation (blip2_mr.py:572-824, interleave branch): an int32 row table
        [B, Le] (>= 0: embedding id, < 0: frame-token row -(idx+1), INT_MIN: zero row) + attention mask."""
        tok = self.t5_tokenizer
        if "add_duration" in self.task:
            video_prompt_end = ["{}<extra_id_0>\n".format(">" + str(round(float(x), 2))) for x in durations]
        ts = torch.as_tensor(timestamps).detach().cpu()
        du = torch.as_tensor(durations).detach().cpu()
        ts_list, dur_list, video_prompt = self._timestamps(ts, du)
        end = tok(video_prompt_end, padding="longest", add_special_tokens=False, truncation=True,
                  max_length=self.max_txt_len, return_tensors="pt")
        if "no_task_prompt" in self.task:
            text_prompt = [q for q in query_prompt]
        else:
            text_prompt = [q + t for q, t in zip(query_prompt, task_prompt)]
        text = tok(text_prompt, padding="longest", truncation=True, max_length=self.max_txt_len, return_tensors="pt")
        if not self.interleave_data:
            # blip2_mr.py:784-822: [video_prompt (the timestamps as text) | all frame tokens | video_prompt_end | query + task]
            vp = tok(video_prompt, padding="longest", add_special_tokens=False, truncation=True, max_length=self.max_txt_len,
                     return_tensors="pt")
            frame_rows = -(np.arange(B)[:, None] * (T * n) + np.arange(T * n)[None, :]) - 1
            table = np.concatenate([vp.input_ids.numpy(), frame_rows, end.input_ids.numpy(), text.input_ids.numpy()], axis=1)
            atts = torch.cat([vp.attention_mask, torch.ones((B, T * n), dtype=torch.long), end.attention_mask, text.attention_mask], dim=1)
            return table.astype(np.int32), atts, [p_ + "frames" + e_ for p_, e_ in zip(video_prompt, video_prompt_end)]
        rows = []
        for j in range(B):
            ts_ids = self._clean_ids(ts_list[j].tolist())
            dur = dur_list[j]
            dur_ids = self._clean_ids([dur.item() if torch.is_tensor(dur) else dur])[0]
            seq = []
            base = j * T * n
            for i in range(T):
                seq.extend(range(-(base + i * n) - 1, -(base + (i + 1) * n) - 1, -1))
                seq.extend(ts_ids[i])
            seq.append(self.seperator_token)
            seq.extend(dur_ids)
            rows.append(seq)
        Lv = max(len(r) for r in rows)
        n_end, n_text = end.input_ids.shape[1], text.input_ids.shape[1]
        table = np.empty((B, Lv + n_end + n_text), dtype=np.int64)
        for j, r in enumerate(rows):
            table[j, :Lv - len(r)] = INT_MIN                  # left padding = pad_token_id * ones = zero rows (:744-754)
            table[j, Lv - len(r):Lv] = r
        table[:, Lv:Lv + n_end] = end.input_ids.numpy()
        table[:, Lv + n_end:] = text.input_ids.numpy()
        atts = torch.cat([torch.ones((B, Lv), dtype=torch.long), end.attention_mask, text.attention_mask], dim=1)
        return table.astype(np.int32), atts, video_prompt

    def prompt_concatenation(self, timestamps, durations, frames_for_t5, frames_atts_for_t5, video_prompt_end,
                             query_prompt, task_prompt):
        """blip2_mr.py:572-824.  The host builds only the row table; one gather kernel writes inputs_embeds.
        -> (inputs_embeds fp32 [B, L, D] cuda, attention_mask long [B, L] cuda, video_prompt list[str])."""
        B, TN, C = frames_for_t5.shape
        n = 1 if self.frame_token_aggregation else self.num_query_token
        table, atts, video_prompt = self.build_prompt_table(timestamps, durations, B, TN // n, n, video_prompt_end,
                                                            query_prompt, task_prompt)
        Le = table.shape[1]
        idx = torch.from_numpy(table).to("cuda", non_blocking=True)
        _, _, t5 = self.engines()
        inputs = torch.empty((B * Le, C), dtype=torch.float32, device="cuda")
        ops.gather_rows(idx.reshape(-1), t5.emb, frames_for_t5.reshape(B * TN, C), inputs)
        return inputs.view(B, Le, C), atts.to("cuda", non_blocking=True), video_prompt

    # ---------------------------------------------------------------------------------------------
    def forward(self, samples):
        if "QA" in self.task:                                # blip2_mr.py:300-307
            return self.forward_QA(samples)
        return self.forward_mr(samples)

    # ---- two-stage video QA (blip2_mr.py:309-431, 990-1099, 1233-1314) --------------------------
    def _qa_relevant_frames(self, samples, generate_kwargs=None):
        """Stage 1: the clip's window -- proposed by the localizer (generate), the ground truth (oracle setting) or the whole
        video -- and `num_frames_for_answer` of the sampled frames inside it.  -> (relevant_moments, frames [b, n, 3, H, W])."""
        n = self.num_frames_for_answer
        if self.use_localizer:
            out_mr = self.generate(samples, **(generate_kwargs or {}))
            if self.resample_frames:                         # re-decode inside the proposed window (blip2_mr.py:340-348)
                return qa.relevant_frames_resampled(samples, out_mr["prediction"], self.video_processor_answerer_eval)
            moments = qa.relevant_moments_from_predictions(out_mr["prediction"], samples["duration"])
        elif self.use_oracle_localizer and generate_kwargs is not None:         # videoQA_generate only (blip2_mr.py:1057-1075)
            rw = samples["relevant_windows"]
            moments = [m[0] for m in (rw.tolist() if torch.is_tensor(rw) else rw)]
            if self.resample_frames:
                return qa.relevant_frames_resampled(samples, moments, self.video_processor_answerer_eval)
        else:
            moments = [[0, d.item() if torch.is_tensor(d) else d] for d in samples["duration"]]
        return moments, qa.extract_frames(samples, moments, n)

    def _qa_inputs(self, frames, texts):
        """[frame tokens ; embed_tokens(question)] (blip2_mr.py:374-393) through the interleave-gather kernel: a row table of
        frame rows (negative) followed by the question's token ids.  -> (inputs fp32 [B, L, D], kmask int32 [B, L])."""
        ans = self._answerer
        B, TN, C = frames.shape
        tok = self.t5_tokenizer(texts, padding="longest", truncation=True, max_length=self.max_txt_len, return_tensors="pt")
        ids = tok.input_ids.numpy().astype(np.int64)
        table = np.concatenate([-(np.arange(B)[:, None] * TN + np.arange(TN)[None, :]) - 1, ids], axis=1).astype(np.int32)
        kmask = np.concatenate([np.ones((B, TN), np.int32), tok.attention_mask.numpy().astype(np.int32)], axis=1)
        Le = table.shape[1]
        idx = torch.from_numpy(np.ascontiguousarray(table.reshape(-1))).to("cuda", non_blocking=True)
        inputs = torch.empty((B * Le, C), dtype=torch.float32, device="cuda")
        ops.gather_rows(idx, ans.emb, frames.reshape(B * TN, C), inputs)
        return inputs.view(B, Le, C), torch.from_numpy(np.ascontiguousarray(kmask)).to("cuda", non_blocking=True)

    def forward_QA(self, samples, want_logits=False):
        """blip2_mr.py:309-431: stage 1 without gradients (localizer / uniform window, frame selection, ViT -> Q-Former ->
        t5_proj), stage 2 the answerer's teacher-forced loss; only the answerer's LoRA adapters receive gradients."""
        self.engines()
        samples["relevant_windows"] = [[0, 0]]               # dummy answer (blip2_mr.py:314)
        samples["query_id"] = samples["question_id"]
        need_grad = self.training and torch.is_grad_enabled()
        with torch.no_grad():
            moments, rel = self._qa_relevant_frames(samples)
            samples["relevant_frames"] = rel
            # the localizer ran the eval-mode arithmetic (generate); stage 2 follows module.training
            self._set_dropout(self.training)
            frames, _ = self.get_frame_embeddings_and_attentions(rel)
        ans = self._answerer
        inputs, kmask = self._qa_inputs(frames, samples["qa_input"])
        out_tok = self.t5_tokenizer(samples["qa_output"], padding="longest", truncation=True, max_length=self.max_txt_len,
                                    return_tensors="pt")
        labels = out_tok.input_ids.masked_fill(out_tok.input_ids == self.t5_tokenizer.pad_token_id, -100)
        if need_grad:
            self._gflat.zero_()
        out = ans.loss(inputs, kmask, labels, out_tok.attention_mask, backward=need_grad, want_logits=want_logits)
        loss = out["loss"].reshape(())
        if need_grad:
            loss = _HandOverGrads.apply(loss, self, *self._grad_params)
        res = {"loss": loss}
        if want_logits:
            res.update(logits=out["logits"], relevant_moments=moments, labels=labels, inputs_embeds=inputs)
        return res

    @torch.no_grad()
    def videoQA_answer(self, samples, max_length=50, min_length=8, **unused):
        """blip2_mr.py:1233-1314: greedy decode of the answerer; the answer is the arg-max over the five letter ids of the scores
        of the SECOND generated position (`outputs_qa.scores[1]`).  Greedy search with min_length only suppresses eos before
        that position, so two decode steps give exactly those scores."""
        self.engines()
        self._set_dropout(False)
        ans = self._answerer
        frames, _ = self.get_frame_embeddings_and_attentions(samples["relevant_frames"])
        inputs, kmask = self._qa_inputs(frames, samples["qa_input"])
        B, Le, _ = inputs.shape
        enc_ext, kmask = ans.encode(inputs, kmask)
        st = ans.init_decode(enc_ext, B, Le, 1, 2)
        eos = self.t5_tokenizer.eos_token_id
        tok0 = torch.full((B,), self.t5_tokenizer.pad_token_id, dtype=torch.long, device="cuda")      # decoder_start_token_id
        l0 = ans.decode_step(st, tok0, 0, kmask).clone()
        if min_length > 1:
            l0[:, eos] = float("-inf")                       # MinLengthLogitsProcessor
        l1 = ans.decode_step(st, l0.argmax(-1), 1, kmask)
        pred = l1[:, qa.ANSWER_IDS].argmax(-1).cpu().tolist()
        return {"output_text": pred, "answer": samples["qa_output"], "qid": samples["question_id"],
                "relevant_moments_gt": samples["relevant_windows"], "answer_scores": l1[:, qa.ANSWER_IDS].float().cpu()}

    @torch.no_grad()
    def videoQA_generate(self, samples, num_frames_for_answer=4, use_nucleus_sampling=False, num_beams=5, max_length=50,
                         min_length=8, top_p=0.9, repetition_penalty=1.0, length_penalty=1.0, num_captions=1, temperature=1,
                         output_attentions=False):
        """blip2_mr.py:990-1099."""
        if "relevant_windows" not in samples:
            samples["relevant_windows"] = [[0, 0]]
        samples["query_id"] = samples["question_id"]
        moments, rel = self._qa_relevant_frames(samples, dict(use_nucleus_sampling=use_nucleus_sampling, num_beams=num_beams,
                                                              max_length=max_length, min_length=min_length,
                                                              length_penalty=length_penalty))
        samples["relevant_frames"] = rel
        out = self.videoQA_answer(samples)
        out["relevant_moments"] = [moments]
        return out

    # ---- gradient hand-over ---------------------------------------------------------------------
    def _grad_views(self, flat):
        out, off = [], 0
        for p in self._grad_params:
            out.append(flat[off:off + p.numel()].view(p.shape))
            off += p.numel()
        return out

    def _hand_over(self, g):
        """g * (flat gradients of this step) -> per-parameter views.  The views alias self._hflat when no gradient is
        pending (optimizer.zero_grad(set_to_none=True), the torch default); under gradient accumulation a fresh buffer is
        used so that earlier .grad tensors (which may alias _hflat) are not overwritten."""
        ps = self._grad_params
        pending = any(p.grad is not None for p in (ps[0], ps[len(ps) // 2], ps[-1]))
        dst = torch.empty_like(self._gflat) if pending else self._hflat
        torch.mul(self._gflat, g.to(self._gflat.dtype), out=dst)
        return self._grad_views(dst)

    def flat_grads(self):
        """The flat buffer that currently backs every trainable .grad (for a single all-reduce), or None."""
        ps = getattr(self, "_grad_params", None)
        if not ps or any(p.grad is None for p in ps):
            return None
        base, off = self._hflat.data_ptr(), 0
        for p in ps:
            if p.grad.data_ptr() != base + 4 * off or not p.grad.is_contiguous():
                return None
            off += p.numel()
        return self._hflat

    # ---- host half of a step --------------------------------------------------------------------
    @staticmethod
    def _video_of(samples):
        image = samples["video"]
        return torch.stack(image) if isinstance(image, list) else image

    def _host_phase(self, samples, bucket=None):
        """Everything forward_mr does on the host (blip2_mr.py:513-534): timestamps -> strings -> ids, the interleave row
        table, the tokenised answer.  -> numpy int32 arrays (idx [B*Le], kmask [B,Le], labels/dec_ids/dmask [B,Ld]).
        bucket = (e, d): pad Le / Ld up to multiples with masked pad tokens / ignored targets (exact: masked keys and
        -100 targets contribute nothing to the loss or to any gradient)."""
        b, t = self._video_of(samples).shape[:2]
        n = 1 if self.frame_token_aggregation else self.num_query_token
        table, atts, _ = self.build_prompt_table(samples["timestamps"], samples["duration"], b, t, n,
                                                 samples["video_prompt_end"], samples["query_prompt"], samples["task_prompt"])
        ans = self.t5_tokenizer(samples["relevant_windows"], padding="longest", truncation=True,
                                max_length=self.max_txt_len, return_tensors="pt")
        labels = ans.input_ids.masked_fill(ans.input_ids == self.t5_tokenizer.pad_token_id, -100).numpy().astype(np.int32)
        dmask = ans.attention_mask.numpy().astype(np.int32)
        kmask = atts.numpy().astype(np.int32)
        if bucket:
            Le, Ld = table.shape[1], labels.shape[1]
            pe, pd = -Le % bucket[0], -Ld % bucket[1]
            if pe:
                table = np.concatenate([table, np.full((b, pe), self.pad_token_id, np.int32)], axis=1)
                kmask = np.concatenate([kmask, np.zeros((b, pe), np.int32)], axis=1)
            if pd:
                labels = np.concatenate([labels, np.full((b, pd), -100, np.int32)], axis=1)
                dmask = np.concatenate([dmask, np.zeros((b, pd), np.int32)], axis=1)
        dec_ids = np.zeros_like(labels)                      # _shift_right, modeling_t5.py:919-948
        dec_ids[:, 1:] = labels[:, :-1]
        dec_ids[dec_ids == -100] = 0
        return dict(idx=np.ascontiguousarray(table.reshape(-1)), kmask=np.ascontiguousarray(kmask),
                    labels=np.ascontiguousarray(labels), dec_ids=dec_ids, dmask=np.ascontiguousarray(dmask),
                    b=b, t=t, Le=table.shape[1], Ld=labels.shape[1])

    # ---- device half of a step (no host synchronisation: capturable) ------------------------------
    def _device_phase(self, video, idx, kmask, labels, dec_ids, dmask, need_grad, want_logits=False, loss_out=None, front=None):
        """ViT -> ln_vision -> Q-Former -> t5_proj -> interleave gather -> T5 loss (-> backward into the flat gradient
        buffer).  idx int32 [B*Le], kmask/dmask int32, labels/dec_ids int64 -- all on the GPU."""
        vit, qf, t5 = self.engines()
        d = self.dims
        b, t = video.shape[:2]
        if front is not None:
            frames, qh, qh16 = front
        else:
            frames, _, (qh, qh16) = self.get_frame_embeddings_and_attentions(video, want_aux=True)
        B, TN, C = frames.shape
        Le = idx.numel() // B
        inputs = torch.empty((B * Le, C), dtype=torch.float32, device="cuda")
        ops.gather_rows(idx, t5.emb, frames.reshape(B * TN, C), inputs)
        if need_grad:
            self._gflat.zero_()
        out = t5.loss_device(inputs.view(B, Le, C), kmask, labels, dec_ids, dmask, backward=need_grad,
                             want_logits=want_logits, loss_out=loss_out)
        if need_grad:
            n = d.num_query
            M = b * t * n
            d_frames = torch.zeros((b * t * (1 if self.frame_token_aggregation else n), d.d_model),
                                   dtype=torch.float32, device="cuda")
            din = out["d_inputs_embeds"].reshape(-1, d.d_model)
            ops.scatter_frames(idx, din, d_frames)
            if self.frame_token_aggregation:
                full = torch.empty((M, d.d_model), dtype=torch.float32, device="cuda")
                ops.group_mean_bwd(d_frames, full, b * t, n, d.d_model)
                d_frames = full
            self._t5_proj_grads(d_frames, qh, M)
        if want_logits:
            out.update(inputs_embeds=inputs.view(B, Le, C), qformer=qh.view(b * t, d.num_query, -1), frames_for_t5=frames)
        return out

    def _device_step(self, st):
        """What a CUDA graph holds: LoRA / t5_proj re-pack (the optimiser changed them) + forward + backward."""
        self.engines()
        self._in_device_step = True
        try:
            _, qf, t5 = self._engines
            t5.refresh()
            if st.front is None:
                qf.set_t5_proj(self.t5_proj.weight, self.t5_proj.bias)
            x = st.d
            self._device_phase(st.video, x["idx"], x["kmask"], x["labels"].to(torch.int64), x["dec_ids"].to(torch.int64),
                               x["dmask"], need_grad=True, loss_out=st.loss, front=st.front.out if st.front is not None else None)
        finally:
            self._in_device_step = False

    def _front_step(self, fr):
        """What the front graph holds: t5_proj re-cast (the optimiser changed it) + ViT -> ln_vision -> Q-Former -> t5_proj."""
        self.engines()
        self._in_device_step = True
        try:
            _, qf, _ = self._engines
            qf.set_t5_proj(self.t5_proj.weight, self.t5_proj.bias)
            frames, _, (qh, qh16) = self.get_frame_embeddings_and_attentions(fr.video, want_aux=True)
            fr.out = (frames, qh, qh16)
        finally:
            self._in_device_step = False

    def _run_front(self, video, vdt):
        """Copy the frames into the front graph's buffer and launch it (eager on first sight, captured on the second)."""
        b, t = video.shape[:2]
        fkey = (b, t, bool(self.frame_token_aggregation), vdt, self._engines[1].drop is not None)
        fr = self._fronts.get(fkey)
        if fr is None:
            fr = self._fronts[fkey] = _GraphedFront(b, t, self.dims.img_size, vdt)
        fr.video.copy_(video.reshape(fr.video.shape), non_blocking=True)
        seen = self._seen[("front",) + fkey] = self._seen.get(("front",) + fkey, 0) + 1
        if fr.graph is not None:
            fr.graph.replay()
            _lib.launch_count += fr.n_launch
        elif seen < 2:
            self._front_step(fr)
        else:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count
            with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                self._front_step(fr)
            fr.n_launch = _lib.launch_count - l0
            if self._graph_pool is None:
                self._graph_pool = g.pool()
            fr.graph = g
            g.replay()
        return fr

    def _graphed_step(self, samples):
        video = self._video_of(samples)
        vdt = torch.uint8 if video.dtype == torch.uint8 else torch.float32
        front = None
        if self.split_graph:
            self.engines()
            front = self._run_front(video, vdt)              # the GPU starts on the ViT; the prompt table is built under it
        host = self._host_phase(samples, bucket=self.graph_bucket)
        key = (host["b"], host["t"], host["Le"], host["Ld"], bool(self.frame_token_aggregation), vdt,
               self._engines[2].drop is not None if self._engines is not None else None, front is not None)
        st = self._steps.pop(key, None)
        if st is not None and st.front is not front:         # captured against another front's output buffers
            st = None
        if st is None:
            while len(self._steps) >= self.max_graphs:       # least recently used shape goes first
                self._steps.pop(next(iter(self._steps)))
            if not any(x.graph is not None for x in self._steps.values()) and not any(f.graph is not None for f in self._fronts.values()):
                self._graph_pool = None                      # no live graph holds the old pool any more
            st = _GraphedStep(host["b"], host["t"], host["Le"], host["Ld"], self.dims.img_size, vdt, front=front)
        self._steps[key] = st
        self.engines()
        st.stage(host, None if front is not None else video)
        seen = self._seen[key] = self._seen.get(key, 0) + 1
        if st.graph is not None:
            st.graph.replay()
            _lib.launch_count += st.n_launch                 # the kernels the replay just ran (bench.py: gpu_launches)
        elif seen < 2:
            self._device_step(st)                            # first sight of this shape: eager (also the warm-up)
        else:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count
            with torch.cuda.graph(g, pool=self._graph_pool, capture_error_mode="thread_local"):
                self._device_step(st)
            st.n_launch = _lib.launch_count - l0             # C-ABI kernels recorded into the graph
            if self._graph_pool is None:
                self._graph_pool = g.pool()
            st.graph = g
            g.replay()
        st.calls += 1
        self._lora_versions = None                           # the step re-packed LoRA itself; eager callers re-check
        return st.loss

    def _set_dropout(self, on, advance=True):
        """Switch the engines' train-mode dropout for the calls that follow; a step that uses it draws a new seed word first
        (outside any graph capture: the captured kernels read the word from device memory)."""
        _, qf, t5 = self.engines()
        st = self.drop_state if (on and self.train_dropout) else None
        qf.drop = t5.drop = st
        if self._answerer is not None:                       # QA: the localizer only ever runs generate (eval arithmetic)
            t5.drop, self._answerer.drop = None, st
        if st is not None and advance:
            st.advance()
        return st is not None

    def forward_mr(self, samples, want_logits=False):
        """blip2_mr.py:433-570."""
        need_grad = self.training and torch.is_grad_enabled()
        graphed = need_grad and self.cuda_graphs and not want_logits
        self._defer_refresh = graphed and self._engines is not None      # the (captured) step re-packs LoRA / t5_proj itself:
        try:                                                              # no second, eager re-pack in front of it
            self.engines()
            self._set_dropout(self.training)                 # nn.Dropout follows module.training, with or without grad
            if graphed:
                loss = self._graphed_step(samples).reshape(())
                return {"loss": _HandOverGrads.apply(loss, self, *self._grad_params)}
        finally:
            self._defer_refresh = False
        host = self._host_phase(samples)
        dev = {k: torch.from_numpy(host[k]).to("cuda", non_blocking=True) for k in ("idx", "kmask", "labels", "dec_ids", "dmask")}
        video = self._video_of(samples)
        video = video.to(device="cuda", non_blocking=True) if video.dtype == torch.uint8 else \
            video.to(device="cuda", dtype=torch.float32, non_blocking=True)
        out = self._device_phase(video, dev["idx"], dev["kmask"], dev["labels"].to(torch.int64), dev["dec_ids"].to(torch.int64),
                                 dev["dmask"], need_grad, want_logits)
        loss = out["loss"].reshape(())
        if need_grad:
            loss = _HandOverGrads.apply(loss, self, *self._grad_params)
        res = {"loss": loss}
        if want_logits:
            res.update(logits=out["logits"], inputs_embeds=out["inputs_embeds"], attention_mask=dev["kmask"].long(),
                       labels=torch.from_numpy(host["labels"]).long(), qformer=out["qformer"], frames_for_t5=out["frames_for_t5"])
        return res

    def _t5_proj_grads(self, d_frames, qh, M):
        """dW = dF^T . h, db = colsum(dF) for t5_proj (trainable, blip2_mr.py:291 note in SURVEY.md §3.1), written into
        their slots of the flat gradient buffer (zeroed at the start of the step)."""
        d = self.dims
        BF = torch.bfloat16
        Mp = (M + 7) // 8 * 8
        df16 = torch.empty((M, d.d_model), dtype=BF, device="cuda")
        ops.cast_to(d_frames, df16)
        h16 = torch.empty((M, d.qf_hidden), dtype=BF, device="cuda")
        ops.cast_to(qh.contiguous(), h16)
        dft = torch.zeros((d.d_model, Mp), dtype=BF, device="cuda")
        ht = torch.zeros((d.qf_hidden, Mp), dtype=BF, device="cuda")
        ops.transpose16(df16, dft, M, d.d_model)
        ops.transpose16(h16, ht, M, d.qf_hidden)
        ops.gemm(dft, ht, out=self._g_projW)                  # [2048, 768]
        ops.colsum(d_frames, self._g_projb)

    # ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def generate(self, samples, use_nucleus_sampling=False, num_beams=5, max_length=50, min_length=1, top_p=0.9,
                 repetition_penalty=1.0, length_penalty=1.0, num_captions=1, temperature=1, output_attentions=False):
        """blip2_mr.py:826-946.  Beam search restates transformers 4.46.1 semantics (see generation.py) with
        cross K/V projected once and an incremental decoder."""
        from .generation import beam_search
        if use_nucleus_sampling:
            raise NotImplementedError("nucleus sampling is not used by the moment-retrieval task (moment_retrieval.py:33-36)")
        _, _, t5 = self.engines()
        self._set_dropout(False)                             # the task calls generate under model.eval() (moment_retrieval.py:33-36)
        frames, frames_atts = self.get_frame_embeddings_and_attentions(samples["video"])
        inputs, atts, video_prompt = self.prompt_concatenation(samples["timestamps"], samples["duration"], frames,
                                                               frames_atts, samples["video_prompt_end"],
                                                               samples["query_prompt"], samples["task_prompt"])
        seqs = beam_search(t5, inputs, atts, num_beams=num_beams, max_new_tokens=max_length, min_length=min_length,
                           length_penalty=length_penalty, eos_id=self.t5_tokenizer.eos_token_id,
                           pad_id=self.t5_tokenizer.pad_token_id)
        pred_ans = self.t5_tokenizer.batch_decode(seqs, skip_special_tokens=True)
        out = {}
        dur = samples["duration"]
        out["duration"] = dur.tolist() if isinstance(dur, torch.Tensor) else dur
        if self.input_time_format in ("relative_integers", "relative_floats"):
            prediction = [self.post_process(p) for p in pred_ans]
            out["prediction"] = mr_utils.convert_to_absolute_time(prediction, out["duration"], self.input_time_format)
        else:
            out["prediction"] = [self.post_process(p) for p in pred_ans]
        out["raw_prediction"] = pred_ans
        out["answer"] = samples["relevant_windows"]
        out["qid"] = samples["query_id"]
        out["sequences"] = seqs
        return out
