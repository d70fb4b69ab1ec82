# bench_reference.jl -- times the UNMODIFIED reference (pxl-th/Monodepth2.jl) on the hot path of this repo:
# everything of `train_loss` after `model(...)` (src/training.jl:29-77), forward + backward through Zygote, on the
# synthetic workload of bench.py (BASELINE.json configs[1]: 416x128, batch 8, C=1, 3 frames, 4 scales, no automask).
#
# Needs a Julia toolchain with the reference's dependencies instantiated (none of which exists in this repo's build
# image, which is why bench.py times a PyTorch-CPU restatement instead -- BASELINE.md section 4):
#
#     julia --project=/path/to/Monodepth2.jl -t auto baseline/julia/bench_reference.jl [cpu|gpu] [steps]
#
# Prints one JSON line shaped like bench.py's reference arm.  NOT executed here (no Julia in the image).
using Monodepth, Flux, Zygote, Statistics, Random, Printf
using Monodepth: TrainCache, Params, SSIM, Backproject, Project, Pose, train_loss

const W, H, N, C, L = 416, 128, 8, 1, 3
const SCALES = [0.125, 0.25, 0.5, 1.0]
device = length(ARGS) >= 1 && ARGS[1] == "gpu" ? gpu : cpu
steps = length(ARGS) >= 2 ? parse(Int, ARGS[2]) : 10
transfer = device ∘ f32

# a stand-in for Model (src/model.jl:8-20) that returns fixed decoder outputs as trainable leaves, so that the timed
# region is the loss path alone (the networks are outside this repo's scope)
struct FixedOutputs{D, P}
    disparities::D
    poses::P
end
Flux.@functor FixedOutputs
(m::FixedOutputs)(x, source_ids, target_id) = (m.disparities, m.poses)

Random.seed!(42)
x = transfer(rand(Float32, W, H, C, L, N))
disps = [transfer(Float32.(0.1 .+ 0.8 .* rand(round(Int, W * s), round(Int, H * s), 1, N))) for s in SCALES]
poses = [Pose(transfer(0.01f0 .* randn(Float32, 3, N)), transfer(0.01f0 .* randn(Float32, 3, 1, N))) for _ in 1:2]
model = FixedOutputs(disps, poses)
f = 0.58 * W
K = Float64[f 0 W/2; 0 f H/2; 0 0 1]
cache = TrainCache(transfer(SSIM()), transfer(Backproject(; width=W, height=H)), transfer(Project(; width=W, height=H)),
                   transfer(K), transfer(inv(K)), 2, [1, 3], SCALES)
params = Params(; batch_size=N, target_size=(W, H), disparity_smoothness=1e-3, automasking=false)
θ = Flux.params(model)

step() = gradient(θ) do
    train_loss(model, x, nothing, cache, params, false)[1]
end

step(); step()                                  # compile + warm up
times = Float64[]
for _ in 1:steps
    t0 = time_ns()
    step()
    device === gpu && Monodepth.CUDA.synchronize()
    push!(times, (time_ns() - t0) / 1e9)
end
@printf("{\"impl\": \"reference (Julia, %s)\", \"metric\": \"train frames/s @416x128 R18: view-synthesis loss path\", \"value\": %.3f, \"unit\": \"frames/s\", \"steps\": %d, \"ms_per_step\": %.2f, \"threads\": %d, \"best\": %.3f}\n",
        string(device), N / mean(times), steps, 1e3 * mean(times), Threads.nthreads(), N / minimum(times))
