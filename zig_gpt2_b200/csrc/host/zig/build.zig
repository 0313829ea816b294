// build.zig -- the reference's build (zig_gpt2 build.zig:6-82) with the BLAS link replaced by the CUDA shim.
// Zig 0.11.  Point -Dshim-dir at the directory holding libzg_b200.so (default ../zig_gpt2_b200) and
// -Dcuda-dir at the CUDA toolkit (default /usr/local/cuda).
const std = @import("std");

fn linkShim(step: *std.Build.Step.Compile, shim_dir: []const u8, include_dir: []const u8) void {
    step.linkLibC();
    step.addIncludePath(.{ .path = include_dir });
    step.addLibraryPath(.{ .path = shim_dir });
    step.addRPath(.{ .path = shim_dir });
    step.linkSystemLibrary("zg_b200"); // libzg_b200.so: kernels + static cudart; needs only libcuda at run time
}

pub fn build(b: *std.Build) void {
    const target = b.standardTargetOptions(.{});
    const optimize = b.standardOptimizeOption(.{});
    const shim_dir = b.option([]const u8, "shim-dir", "directory containing libzg_b200.so") orelse "../zig_gpt2_b200";
    const include_dir = b.option([]const u8, "include-dir", "directory containing zg_b200.h") orelse "../include";

    const exe = b.addExecutable(.{
        .name = "zig_gpt2",
        .root_source_file = .{ .path = "src/main.zig" },
        .target = target,
        .optimize = optimize,
    });
    linkShim(exe, shim_dir, include_dir);
    b.installArtifact(exe);

    const run_cmd = b.addRunArtifact(exe);
    run_cmd.step.dependOn(b.getInstallStep());
    if (b.args) |args| {
        run_cmd.addArgs(args);
    }
    const run_step = b.step("run", "Run the app");
    run_step.dependOn(&run_cmd.step);

    const unit_tests = b.addTest(.{
        .root_source_file = .{ .path = "src/tests.zig" },
        .target = target,
        .optimize = optimize,
    });
    linkShim(unit_tests, shim_dir, include_dir);
    const run_unit_tests = b.addRunArtifact(unit_tests);
    const test_step = b.step("test", "Run unit tests");
    test_step.dependOn(&run_unit_tests.step);
}
