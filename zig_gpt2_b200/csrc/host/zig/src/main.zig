//! main.zig -- the reference's model / generation loop (zig_gpt2 src/main.zig) over the B200 CUDA shim.
//! GPTConfig, State, MLP, Block, GPT.forward / sample, load_* and generate keep their names and signatures;
//! every slice is device memory; GPT.forward runs the fused persistent decode engine.  Zig 0.11, not
//! compilable in the build image -- see zig/README.md and INTEGRATION.md.
const std = @import("std");
const ops = @import("ops.zig");
const bpe = @import("bpe.zig"); // the reference's tokenizer, unchanged: it is CPU code and never touches the shim
const c = ops.c;

pub const GPTConfig = struct {
    const Self = @This();

    vocab_size: usize,
    context_size: usize,
    n_layer: usize,
    n_heads: usize,
    n_embed: usize,

    pub fn init(vocab_size: usize, context_size: usize, n_layer: usize, n_heads: usize, n_embed: usize) Self {
        return Self{ .vocab_size = vocab_size, .context_size = context_size, .n_layer = n_layer, .n_heads = n_heads, .n_embed = n_embed };
    }

    fn raw(self: Self) c.zg_config {
        return c.zg_config{ .vocab_size = self.vocab_size, .context_size = self.context_size, .n_layer = self.n_layer, .n_heads = self.n_heads, .n_embed = self.n_embed };
    }
};

/// The preallocated buffer set shared by all layers: device memory, allocated once.
pub const State = struct {
    const Self = @This();

    raw: c.zg_state,
    pos_emb: []f32,
    x: []f32,
    o: []f32,
    logits: []f32,
    decoded: []u8, // host
    _h: []f32,
    _4xh: []f32,
    _qkv: []f32,
    _q: []f32,
    _attn: []f32,

    pub fn init(config: GPTConfig, allocator: std.mem.Allocator) !Self {
        var raw: c.zg_state = undefined;
        const cfg = config.raw();
        // the reference's _k/_v transposed-cache scratch (2 x context x n_embed floats) is not needed
        if (c.zg_state_init(&raw, &cfg, 0) != 0) return ops.DeviceError.CudaFailure;
        const E = config.n_embed;
        return Self{
            .raw = raw,
            .pos_emb = raw.pos_emb[0..E],
            .x = raw.x[0..E],
            .o = raw.o[0..E],
            .logits = raw.logits[0..config.vocab_size],
            .decoded = try allocator.alloc(u8, 20),
            ._h = raw._h[0..E],
            ._4xh = raw._4xh[0 .. 4 * E],
            ._qkv = raw._qkv[0 .. 3 * E],
            ._q = raw._q[0..E],
            ._attn = raw._attn[0..config.context_size],
        };
    }
};

const MLP = struct {
    const Self = @This();

    c_fc: ops.Linear,
    c_proj: ops.Linear,

    pub fn init(c_fc: ops.Linear, c_proj: ops.Linear) MLP {
        return MLP{ .c_fc = c_fc, .c_proj = c_proj };
    }

    fn raw(self: Self) c.zg_mlp {
        return c.zg_mlp{ .c_fc = self.c_fc.raw(), .c_proj = self.c_proj.raw() };
    }

    /// Result in state.o.
    pub fn forward(self: Self, inputs: []const f32, state: State) void {
        const m = self.raw();
        c.zg_mlp_forward(&m, inputs.ptr, inputs.len, &state.raw);
    }
};

const Block = struct {
    const Self = @This();

    n_embed: usize,
    ln_1: ops.LayerNorm,
    attn: ops.CausalSelfAttention,
    ln_2: ops.LayerNorm,
    mlp: MLP,
    k_cache: []f32, // device, [context_size, n_embed] time-major
    v_cache: []f32,

    pub fn init(n_embed: usize, ln_1: ops.LayerNorm, attn: ops.CausalSelfAttention, ln_2: ops.LayerNorm, mlp: MLP, k_cache: []f32, v_cache: []f32) Self {
        return Self{ .n_embed = n_embed, .ln_1 = ln_1, .attn = attn, .ln_2 = ln_2, .mlp = mlp, .k_cache = k_cache, .v_cache = v_cache };
    }

    pub fn raw(self: Self) c.zg_block {
        return c.zg_block{
            .n_embed = self.n_embed,
            .ln_1 = self.ln_1.raw(),
            .attn = self.attn.raw(),
            .ln_2 = self.ln_2.raw(),
            .mlp = self.mlp.raw(),
            .k_cache = self.k_cache.ptr,
            .v_cache = self.v_cache.ptr,
        };
    }

    /// Result in both state.x and state.o, as in the reference.
    pub fn forward(self: Self, seq_len: usize, inputs: []const f32, state: State) void {
        const b = self.raw();
        c.zg_block_forward(&b, seq_len, inputs.ptr, &state.raw);
    }
};

const GPT = struct {
    const Self = @This();

    config: GPTConfig,
    wte: ops.Embedding,
    wpe: ops.Embedding,
    h: []const Block,
    ln_f: ops.LayerNorm,
    lm_head: ops.Linear,
    raw_blocks: []c.zg_block, // the same blocks in the shim's layout (start-up only)
    engine: *c.zg_engine, // fused persistent decode engine, created once

    pub fn init(config: GPTConfig, wte: ops.Embedding, wpe: ops.Embedding, h: []const Block, ln_f: ops.LayerNorm, lm_head: ops.Linear, state: State, allocator: std.mem.Allocator) !Self {
        var raw_blocks = try allocator.alloc(c.zg_block, h.len);
        for (0..h.len) |i| {
            raw_blocks[i] = h[i].raw();
        }
        const g = c.zg_gpt{
            .config = config.raw(),
            .wte = c.zg_embedding{ .emb_dim = wte.emb_dim, .weight = wte.weight.ptr },
            .wpe = c.zg_embedding{ .emb_dim = wpe.emb_dim, .weight = wpe.weight.ptr },
            .h = raw_blocks.ptr,
            .ln_f = ln_f.raw(),
            .lm_head = lm_head.raw(),
        };
        const engine = c.zg_engine_create(&g, &state.raw) orelse return ops.DeviceError.CudaFailure;
        return Self{ .config = config, .wte = wte, .wpe = wpe, .h = h, .ln_f = ln_f, .lm_head = lm_head, .raw_blocks = raw_blocks, .engine = engine };
    }

    /// Logits land in state.logits (device).  One persistent-kernel launch; nothing is allocated.
    pub fn forward(self: Self, seq_len: usize, token: usize, compute_logits: bool, state: State) void {
        _ = state;
        c.zg_engine_forward(self.engine, seq_len, token, @intFromBool(compute_logits));
    }

    /// Temperature sampling.  The reference seeds a PRNG from wall-clock seconds on every call; here the
    /// uniform draw comes from a caller-owned generator so a run can be reproduced.
    pub fn sample(self: Self, seq_len: usize, temp: f32, token: usize, state: State, random: std.rand.Random) usize {
        _ = state;
        return c.zg_engine_sample(self.engine, seq_len, temp, token, random.float(f64));
    }

    pub fn sample_greedy(self: Self, seq_len: usize, token: usize) usize {
        return c.zg_engine_sample_greedy(self.engine, seq_len, token);
    }
};

fn model_path(allocator: std.mem.Allocator, model_dir: []const u8, name: []const u8, suffix: []const u8) ![]u8 {
    return std.fmt.allocPrint(allocator, "{s}/raw/model-{s}{s}", .{ model_dir, name, suffix });
}

pub fn load_linear(name: []const u8, in_features: usize, out_features: usize, model_dir: []const u8, allocator: std.mem.Allocator) !ops.Linear {
    const weight_path = try model_path(allocator, model_dir, name, "-w");
    defer allocator.free(weight_path);
    var weight = try ops.load_tensor(weight_path, &[_]usize{ in_features, out_features }, f32, allocator);
    const bias_path = try model_path(allocator, model_dir, name, "-b");
    defer allocator.free(bias_path);
    var bias = try ops.load_tensor(bias_path, &[_]usize{out_features}, f32, allocator);
    return ops.Linear.init(in_features, out_features, weight, bias);
}

pub fn load_layer_norm(name: []const u8, n_features: usize, model_dir: []const u8, allocator: std.mem.Allocator) !ops.LayerNorm {
    const weight_path = try model_path(allocator, model_dir, name, "-g");
    defer allocator.free(weight_path);
    var weight = try ops.load_tensor(weight_path, &[_]usize{n_features}, f32, allocator);
    const bias_path = try model_path(allocator, model_dir, name, "-b");
    defer allocator.free(bias_path);
    var bias = try ops.load_tensor(bias_path, &[_]usize{n_features}, f32, allocator);
    return ops.LayerNorm.init(n_features, weight, bias);
}

pub fn load_embedding(name: []const u8, vocab_size: usize, emb_dim: usize, model_dir: []const u8, allocator: std.mem.Allocator) !ops.Embedding {
    const path = try model_path(allocator, model_dir, name, "");
    defer allocator.free(path);
    var weight = try ops.load_tensor(path, &[_]usize{ vocab_size, emb_dim }, f32, allocator);
    return ops.Embedding.init(emb_dim, weight);
}

pub fn load_block(layer_idx: usize, config: GPTConfig, model_dir: []const u8, allocator: std.mem.Allocator) !Block {
    var buf: [64]u8 = undefined;
    const E = config.n_embed;
    const ln_1 = try load_layer_norm(try std.fmt.bufPrint(&buf, "h{d}-ln_1", .{layer_idx}), E, model_dir, allocator);
    const c_attn = try load_linear(try std.fmt.bufPrint(&buf, "h{d}-attn-c_attn", .{layer_idx}), E, 3 * E, model_dir, allocator);
    const c_proj = try load_linear(try std.fmt.bufPrint(&buf, "h{d}-attn-c_proj", .{layer_idx}), E, E, model_dir, allocator);
    const ln_2 = try load_layer_norm(try std.fmt.bufPrint(&buf, "h{d}-ln_2", .{layer_idx}), E, model_dir, allocator);
    const c_fc = try load_linear(try std.fmt.bufPrint(&buf, "h{d}-mlp-c_fc", .{layer_idx}), E, 4 * E, model_dir, allocator);
    const mlp_c_proj = try load_linear(try std.fmt.bufPrint(&buf, "h{d}-mlp-c_proj", .{layer_idx}), 4 * E, E, model_dir, allocator);

    const attn = ops.CausalSelfAttention.init(config.n_heads, E, c_attn, c_proj);
    const mlp = MLP.init(c_fc, mlp_c_proj);
    const k_cache = try ops.alloc(f32, config.context_size * E);
    const v_cache = try ops.alloc(f32, config.context_size * E);
    return Block.init(E, ln_1, attn, ln_2, mlp, k_cache, v_cache);
}

pub fn load_gpt(config: GPTConfig, model_dir: []const u8, state: State, allocator: std.mem.Allocator) !GPT {
    var wte = try load_embedding("wte", config.vocab_size, config.n_embed, model_dir, allocator);
    const wpe = try load_embedding("wpe", config.context_size, config.n_embed, model_dir, allocator);
    var h = try allocator.alloc(Block, config.n_layer);
    for (0..h.len) |i| {
        h[i] = try load_block(i, config, model_dir, allocator);
    }
    const ln_f = try load_layer_norm("ln_f", config.n_embed, model_dir, allocator);
    const lm_head = ops.Linear.init(config.n_embed, config.vocab_size, wte.weight, null); // tied to wte, no bias
    return GPT.init(config, wte, wpe, h, ln_f, lm_head, state, allocator);
}

pub fn load_encoder(model_dir: []const u8, allocator: std.mem.Allocator) !bpe.Encoder {
    const enc_path = try std.fmt.allocPrint(allocator, "{s}/encoder.json", .{model_dir});
    const byte_path = try std.fmt.allocPrint(allocator, "{s}/byte_encoder.json", .{model_dir});
    const parsed_encoder = try ops.load_json(enc_path, allocator);
    const parsed_bytes_encoder = try ops.load_json(byte_path, allocator);
    return bpe.Encoder.init(parsed_encoder.object, parsed_bytes_encoder.object, allocator);
}

/// The reference's loop: prompt tokens are forwarded one at a time without logits, then tokens are sampled
/// up to context_size.  As in the reference, the first sampled step forwards the last prompt token again.
pub fn generate(gpt: GPT, encoder: bpe.Encoder, temp: f32, inputs: []usize, state: State, random: std.rand.Random) void {
    var token: usize = undefined;
    for (0..gpt.config.context_size) |s| {
        if (s < inputs.len) {
            token = inputs[s];
            gpt.forward(s + 1, token, false, state);
        } else {
            token = gpt.sample(s + 1, temp, token, state, random);
        }
        const decoded_len = encoder.decode(&[_]usize{token}, state.decoded);
        std.debug.print("{s}", .{state.decoded[0..decoded_len]});
    }
}

/// Greedy variant: the whole loop is one persistent-kernel launch; tokens stream into a pinned host ring.
/// generate with GPT.sample run entirely on the device: the draw of step s is Philox4x32-10(seed; s, sequence)
/// (zg_philox_uniform), so a run is reproducible and needs no host round trip per token.
pub fn generate_sample(gpt: GPT, encoder: bpe.Encoder, temp: f32, seed: u64, inputs: []usize, out_tokens: []usize, state: State) !void {
    if (c.zg_engine_generate_sample(gpt.engine, inputs.ptr, inputs.len, out_tokens.len, temp, seed, 0, out_tokens.ptr) != 0) return ops.DeviceError.CudaFailure;
    for (out_tokens) |token| {
        const decoded_len = encoder.decode(&[_]usize{token}, state.decoded);
        std.debug.print("{s}", .{state.decoded[0..decoded_len]});
    }
}

pub fn generate_greedy(gpt: GPT, encoder: bpe.Encoder, inputs: []usize, out_tokens: []usize, state: State) !void {
    if (c.zg_engine_generate_greedy(gpt.engine, inputs.ptr, inputs.len, out_tokens.len, out_tokens.ptr) != 0) return ops.DeviceError.CudaFailure;
    for (out_tokens) |token| {
        const decoded_len = encoder.decode(&[_]usize{token}, state.decoded);
        std.debug.print("{s}", .{state.decoded[0..decoded_len]});
    }
}

pub fn main() !void {
    const temp = 0.8;
    const config = GPTConfig.init(50257, 1024, 12, 12, 768);
    const model_dir = "models/124M";

    var gpa = std.heap.GeneralPurposeAllocator(.{}){};
    var arena = std.heap.ArenaAllocator.init(gpa.allocator());
    defer arena.deinit();
    const allocator = arena.allocator();

    if (c.zg_init(0) != 0) return ops.DeviceError.CudaFailure;
    defer _ = c.zg_shutdown();

    var inputs = try allocator.alloc(usize, config.context_size);
    var encoder = try load_encoder(model_dir, allocator);
    defer encoder.deinit();
    var state = try State.init(config, allocator);
    const gpt = try load_gpt(config, model_dir, state, allocator);

    const args = try std.process.argsAlloc(allocator);
    defer std.process.argsFree(allocator, args);
    if (args.len < 2) return error.MissingPrompt;
    const prompt = args[1];

    var prng = std.rand.DefaultPrng.init(@intCast(std.time.timestamp()));
    const input_tokens = encoder.encode(prompt, inputs);
    generate(gpt, encoder, temp, inputs[0..input_tokens], state, prng.random());
}
