//! ops.zig -- the reference's operator surface (zig_gpt2 src/ops.zig:4-307) over the B200 CUDA shim.
//!
//! Every `[]f32` / `[]const f32` here wraps a DEVICE pointer obtained from `alloc` below; the host never
//! dereferences it.  Each `forward` is one call into libzg_b200 (include/zg_b200.h).  Names, argument
//! order and shape semantics are the reference's.  Written for Zig 0.11 like the reference; it cannot be
//! compiled in the build image (no Zig toolchain) -- the tested twins are zig_gpt2_b200/csrc/host/gpt2.hpp
//! (C++) and zig_gpt2_b200/{ops,gpt}.py (ctypes), which make exactly these calls.
const std = @import("std");
pub const c = @cImport(@cInclude("zg_b200.h"));

pub const DeviceError = error{ CudaFailure, OutOfDeviceMemory };

/// Device allocation standing in for `allocator.alloc(f32, n)` (start-up only).
pub fn alloc(comptime T: type, n: usize) DeviceError![]T {
    const raw = c.zg_alloc(n * @sizeOf(T)) orelse return DeviceError.OutOfDeviceMemory;
    const ptr: [*]T = @ptrCast(@alignCast(raw));
    return ptr[0..n];
}

pub fn upload(comptime T: type, dst: []T, src: []const T) DeviceError!void {
    if (c.zg_upload(dst.ptr, src.ptr, src.len * @sizeOf(T)) != 0) return DeviceError.CudaFailure;
}

pub fn download(comptime T: type, dst: []T, src: []const T) DeviceError!void {
    if (c.zg_download(dst.ptr, src.ptr, dst.len * @sizeOf(T)) != 0) return DeviceError.CudaFailure;
}

pub const Linear = struct {
    const Self = @This();

    in_features: usize,
    out_features: usize,
    weight: []const f32, // device, [out_features, in_features] row-major
    bias: ?[]const f32, // device

    pub fn init(in_features: usize, out_features: usize, weight: []const f32, bias: ?[]const f32) Self {
        return Self{ .in_features = in_features, .out_features = out_features, .weight = weight, .bias = bias };
    }

    fn raw(self: Self) c.zg_linear {
        return c.zg_linear{
            .in_features = self.in_features,
            .out_features = self.out_features,
            .weight = self.weight.ptr,
            .bias = if (self.bias) |b| b.ptr else null,
        };
    }

    pub fn forward(self: Self, inputs: []const f32, outputs: []f32) void {
        const l = self.raw();
        c.zg_linear_forward(&l, inputs.ptr, inputs.len, outputs.ptr);
    }
};

pub const Embedding = struct {
    const Self = @This();

    emb_dim: usize,
    weight: []const f32, // device

    pub fn init(emb_dim: usize, weight: []const f32) Self {
        return Self{ .emb_dim = emb_dim, .weight = weight };
    }

    /// `idxs` is a HOST slice (the reference passes `&[1]usize{token}`); `embeddings` is device memory.
    pub fn forward(self: Self, idxs: []const usize, embeddings: []f32) void {
        const e = c.zg_embedding{ .emb_dim = self.emb_dim, .weight = self.weight.ptr };
        c.zg_embedding_forward(&e, idxs.ptr, idxs.len, embeddings.ptr);
    }
};

pub const LayerNorm = struct {
    const Self = @This();

    n_features: usize,
    weight: []const f32,
    bias: []const f32,
    eps: f32 = 1e-5,

    pub fn init(n_features: usize, weight: []const f32, bias: []const f32) Self {
        return Self{ .n_features = n_features, .weight = weight, .bias = bias };
    }

    pub fn raw(self: Self) c.zg_layer_norm {
        return c.zg_layer_norm{ .n_features = self.n_features, .weight = self.weight.ptr, .bias = self.bias.ptr, .eps = self.eps };
    }

    pub fn forward(self: Self, inputs: []f32) void {
        const l = self.raw();
        c.zg_layer_norm_forward(&l, inputs.ptr, inputs.len);
    }
};

pub const CausalSelfAttention = struct {
    const Self = @This();

    n_heads: usize,
    n_embed: usize,
    head_dim: usize,
    c_attn: Linear,
    c_proj: Linear,

    pub fn init(n_heads: usize, n_embed: usize, c_attn: Linear, c_proj: Linear) Self {
        return Self{ .n_heads = n_heads, .n_embed = n_embed, .head_dim = n_embed / n_heads, .c_attn = c_attn, .c_proj = c_proj };
    }

    pub fn raw(self: Self) c.zg_attention {
        return c.zg_attention{
            .n_heads = self.n_heads,
            .n_embed = self.n_embed,
            .head_dim = self.head_dim,
            .c_attn = self.c_attn.raw(),
            .c_proj = self.c_proj.raw(),
        };
    }

    pub fn forward(
        self: Self,
        seq_len: usize,
        inputs: []const f32,
        k_cache: []f32,
        v_cache: []f32,
        outputs: []f32,
        _qkv: []f32,
        _q: []f32,
        _k: []f32,
        _v: []f32,
        _attn: []f32,
    ) void {
        const a = self.raw();
        c.zg_attention_forward(&a, seq_len, inputs.ptr, k_cache.ptr, v_cache.ptr, outputs.ptr, _qkv.ptr, _q.ptr, _k.ptr, _v.ptr, _attn.ptr);
    }

    pub fn split_qkv(self: Self, seq_len: usize, inputs: []const f32, split_idx: usize, outputs: []f32) void {
        const a = self.raw();
        c.zg_split_qkv(&a, seq_len, inputs.ptr, inputs.len, split_idx, outputs.ptr);
    }

    pub fn transpose(shape: [3]usize, inputs: []const f32, outputs: []f32) void {
        c.zg_transpose(&shape, inputs.ptr, inputs.len, outputs.ptr);
    }
};

pub fn gelu(inputs: []f32) void {
    c.zg_gelu(inputs.ptr, inputs.len);
}

pub fn softmax(inputs: []f32) void {
    c.zg_softmax(inputs.ptr, inputs.len);
}

pub fn scaled_dot_product_attention(
    q: []const f32,
    k: []const f32,
    v: []const f32,
    n_heads: usize,
    seq_len: usize,
    head_dim: usize,
    outputs: []f32,
    _attn: []f32,
) void {
    c.zg_sdpa(q.ptr, k.ptr, k.len, v.ptr, n_heads, seq_len, head_dim, outputs.ptr, _attn.ptr);
}

/// Reads a headerless raw tensor file into a DEVICE slice.  `allocator` only provides the host staging
/// buffer, freed before returning.  Unlike the reference a short read is an error.
pub fn load_tensor(path: []const u8, shape: []const usize, comptime dtype: type, allocator: std.mem.Allocator) ![]dtype {
    var n_elements: usize = 1;
    for (shape) |item| {
        n_elements *= item;
    }
    var staging = try allocator.alloc(dtype, n_elements);
    defer allocator.free(staging);
    const fd = try std.fs.cwd().openFile(path, .{});
    defer fd.close();
    const got = try fd.readAll(std.mem.sliceAsBytes(staging));
    if (got != n_elements * @sizeOf(dtype)) return error.EndOfStream;
    var tensor = try alloc(dtype, n_elements);
    try upload(dtype, tensor, staging);
    return tensor;
}

pub fn load_json(path: []const u8, allocator: std.mem.Allocator) !std.json.Value {
    const fd = try std.fs.cwd().openFile(path, .{});
    defer fd.close();
    const buffer = try fd.readToEndAlloc(allocator, 4 * 1024 * 1024);
    return std.json.parseFromSliceLeaky(std.json.Value, allocator, buffer, .{});
}
