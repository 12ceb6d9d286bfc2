# VahanaB200.jl — thin `ccall` wrapper over include/vahana_b200.h with Vahana's function names.
# UNTESTED in this repository: Julia is not installed in the build image.  It shows the binding a Vahana
# maintainer would add where `create_model` generates per-model methods (src/Simulation.jl:115-207).
module VahanaB200

export create_simulation, add_agents!, add_edges!, finish_init!, apply!, num_agents, all_agents, mapreduce_field,
       finish_simulation!

const LIB = get(ENV, "VAHANA_B200_LIB", joinpath(@__DIR__, "..", "csrc", "build", "libvahana_b200.so"))
const EDGE_REF = 256
const AgentID = UInt64

struct AgentTypeDesc; name::Cstring; size::UInt32; hints::UInt32; end
struct EdgeTypeDesc; name::Cstring; size::UInt32; hints::UInt32; target::Int32; size_hint::UInt64; end
struct ModelDesc
    name::Cstring; n_agent_types::UInt32; agent_types::Ptr{AgentTypeDesc}
    n_edge_types::UInt32; edge_types::Ptr{EdgeTypeDesc}; param_size::UInt32
end

last_error() = unsafe_string(ccall((:vb_last_error, LIB), Cstring, ()))
function check(rc::Cint)
    rc == 0 && return
    rc == 1 && throw(AssertionError(last_error()))     # the reference's @assert / @mayassert
    rc == 2 && throw(ArgumentError(last_error()))
    error("vahana_b200 [$rc]: $(last_error())")
end

mutable struct Simulation
    handle::Ptr{Cvoid}
    agenttypes::Vector{DataType}   # registration order = type ids (src/ModelTypes.jl:84-89)
    edgetypes::Vector{DataType}
end
typeid(sim, T) = findfirst(==(T), sim.agenttypes)
ref(sim, T) = T in sim.agenttypes ? Cint(typeid(sim, T)) : Cint(EDGE_REF + findfirst(==(T), sim.edgetypes) - 1)
refs(sim, ts) = Cint[ref(sim, T) for T in (ts isa Union{Tuple,AbstractVector} ? ts : (ts,))]

"create_simulation(model, params): `hints` are the VB_AGENT_* / VB_EDGE_* bitmasks of include/vahana_b200.h"
function create_simulation(name, agenttypes::Vector{DataType}, agenthints, edgetypes::Vector{DataType}, edgehints, targets, params)
    check(ccall((:vb_init, LIB), Cint, (Cint,), 0))
    anames = [string(nameof(T)) for T in agenttypes]; enames = [string(nameof(T)) for T in edgetypes]
    GC.@preserve anames enames begin
        ads = [AgentTypeDesc(pointer(anames[i]), fieldcount(agenttypes[i]) == 0 ? 0 : sizeof(agenttypes[i]), agenthints[i]) for i in eachindex(agenttypes)]
        eds = [EdgeTypeDesc(pointer(enames[i]), fieldcount(edgetypes[i]) == 0 ? 0 : sizeof(edgetypes[i]), edgehints[i], targets[i], 0) for i in eachindex(edgetypes)]
        md = Ref(ModelDesc(pointer(name), length(ads), pointer(ads), length(eds), pointer(eds), sizeof(params)))
        h = Ref{Ptr{Cvoid}}()
        GC.@preserve ads eds check(ccall((:vb_sim_create, LIB), Cint, (Ref{ModelDesc}, Ref{typeof(params)}, Ref{Ptr{Cvoid}}), md, Ref(params), h))
        return Simulation(h[], agenttypes, edgetypes)
    end
end

function add_agents!(sim::Simulation, agents::Vector{T}) where T          # src/AgentMethods.jl:65-89
    ids = Vector{AgentID}(undef, length(agents))
    check(ccall((:vb_add_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{T}, UInt64, Ptr{AgentID}), sim.handle, typeid(sim, T), agents, length(agents), ids))
    ids
end
function add_edges!(sim::Simulation, from::Vector{AgentID}, to::Vector{AgentID}, states::Vector{T}) where T   # src/EdgeMethods.jl:388-523
    e = findfirst(==(T), sim.edgetypes) - 1
    check(ccall((:vb_add_edges, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{AgentID}, Ptr{AgentID}, Ptr{T}, UInt64), sim.handle, e, from, to,
                fieldcount(T) == 0 ? C_NULL : states, length(to)))
end
finish_init!(sim::Simulation) = check(ccall((:vb_finish_init, LIB), Cint, (Ptr{Cvoid},), sim.handle))       # src/Simulation.jl:403-476

"apply!(sim, transition, call, read, write; add_existing, with_edge, seed): `transition` names a registered CUDA functor set"
function apply!(sim::Simulation, transition::String, call, read, write; add_existing = [], with_edge = nothing, seed = 0)   # src/Simulation.jl:720-821
    c, r, w, a = refs(sim, call), refs(sim, read), refs(sim, write), refs(sim, add_existing)
    we = with_edge === nothing ? Cint(-1) : Cint(findfirst(==(with_edge), sim.edgetypes) - 1)
    check(ccall((:vb_apply, LIB), Cint, (Ptr{Cvoid}, Cstring, Ptr{Cint}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Cint, Ptr{Cint}, Cint, Cint, UInt64),
                sim.handle, transition, c, length(c), r, length(r), w, length(w), a, length(a), we, seed))
    sim
end

function num_agents(sim::Simulation, ::Type{T}) where T                   # src/Agent.jl:324-343
    n = Ref{UInt64}(0)
    check(ccall((:vb_num_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ref{UInt64}), sim.handle, typeid(sim, T), n))
    Int(n[])
end
function all_agents(sim::Simulation, ::Type{T}) where T                   # src/Agent.jl:234-285
    n = Ref{UInt64}(0)
    check(ccall((:vb_all_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{Cvoid}, Ptr{Cvoid}, UInt64, Ref{UInt64}), sim.handle, typeid(sim, T), C_NULL, C_NULL, 0, n))
    out = Vector{T}(undef, n[])
    check(ccall((:vb_all_agents, LIB), Cint, (Ptr{Cvoid}, Cint, Ptr{T}, Ptr{Cvoid}, UInt64, Ref{UInt64}), sim.handle, typeid(sim, T), out, C_NULL, n[], n))
    out
end
"mapreduce(sim, a -> a.field, op, T) for op in (+, *, min, max, &, |)   (src/AgentMethods.jl:533-565)"
function mapreduce_field(sim::Simulation, ::Type{T}, field::Symbol, op) where T
    ops = Dict(+ => 0, * => 1, min => 2, max => 3, & => 4, | => 5)
    i = findfirst(==(field), fieldnames(T)); FT = fieldtype(T, i)
    dts = Dict(Int64 => 0, Float64 => 1, Bool => 2, Int32 => 3, Float32 => 4, UInt8 => 5)
    RT = FT <: AbstractFloat ? Float64 : (FT == Bool && (op == (&) || op == (|)) ? Bool : Int64)
    out = Ref{RT}()
    check(ccall((:vb_mapreduce, LIB), Cint, (Ptr{Cvoid}, Cint, Cint, Cint, Cint, Int64, Cint, Cint, Ptr{Cvoid}, Ref{RT}),
                sim.handle, ref(sim, T), fieldoffset(T, i), dts[FT], 0, 0, ops[op], dts[RT], C_NULL, out))
    out[]
end
finish_simulation!(sim::Simulation) = (ccall((:vb_sim_destroy, LIB), Cint, (Ptr{Cvoid},), sim.handle); sim.handle = C_NULL; nothing)

end # module
