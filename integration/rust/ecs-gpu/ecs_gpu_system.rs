//! The GPU system kind for recs, on top of `recs-gpu-sys` (include/recs_gpu.h).
//!
//! Where it plugs in (reference file:line, martinjonsson01/recs):
//!   * `trait System` (crates/ecs/src/systems/mod.rs:48-62): [`GpuSystem`] implements it.  `component_accesses()` is
//!     what `recs_gpu_system_describe` reports, so `PrecedenceGraph::generate` (crates/scheduler/src/schedule/mod.rs:
//!     116-198) orders GPU and CPU systems with the unchanged `precedence_to` rules.
//!   * `SequentiallyIterable::run` (crates/ecs/src/systems/iteration.rs:9-14): one call = membership query on the
//!     host (the unchanged `get_archetype_indices`), residency, `recs_gpu_run_system`.
//!     `try_as_segment_iterable()` is `None`, so `WorkerPool` takes the single-task path
//!     (crates/scheduler/src/executor/mod.rs:276-284) and the task may run on any worker thread: every entry point of the
//!     library may be called from any thread.
//!   * application decorator, shaped like `GraphicalApplication<App>` (crates/gfx-plugin/src/lib.rs:214-377): [`Gpu`]
//!     adds `with_gpu()` to any `ApplicationBuilder`, [`GpuApplication`] delegates every `Application` method and
//!     injects its systems before `run`.
//!
//! `World`'s fields are crate-private (crates/ecs/src/lib.rs:665-700).  The three accessors this file needs are listed in
//! [`WorldColumns`]; they are the only lines that have to be added INSIDE crates/ecs (next to
//! `World::borrow_component_vecs`, lib.rs:807-833).
//!
//! NOT COMPILED in the repository this file ships in (no Rust toolchain there): tests/test_rust_bindings.py checks that
//! every `recs_gpu_*` symbol, constant and struct field used below exists in recs-gpu-sys with the spelled arity.

use std::any::TypeId;
use std::collections::HashMap;
use std::ffi::{c_void, CStr, CString};
use std::fmt::Debug;
use std::sync::{Arc, Mutex};

use crossbeam::channel::Receiver;
use ecs::systems::command_buffers::{CommandBuffer, CommandReceiver};
use ecs::systems::iteration::{SegmentIterable, SequentiallyIterable};
use ecs::systems::{
    ComponentAccessDescriptor, IntoSystem, System, SystemError, SystemParameters, SystemResult,
};
use ecs::{
    Application, ApplicationBuilder, ArchetypeIndex, Entity, Executor, IntoBoxedComponentIter, Schedule, World,
};
use recs_gpu_sys as sys;

// ---------------------------------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------------------------------

/// Owns a `recs_gpu_ctx`.  The library serialises calls per context and selects its device on every call, so the
/// handle is `Send + Sync` (the reference runs a system on an arbitrary worker each tick, worker.rs:143-154).
#[derive(Debug)]
pub struct GpuContext {
    raw: *mut sys::recs_gpu_ctx,
    /// (archetype, component) -> what the device copy mirrors: the host Vec's base pointer, length and the archetype's
    /// structural version at bind time, plus which side holds the newest bytes.
    residency: Mutex<HashMap<(u32, u64), Residency>>,
}

unsafe impl Send for GpuContext {}
unsafe impl Sync for GpuContext {}

#[derive(Debug, Clone, Copy)]
struct Residency {
    host_ptr: usize,
    len: u64,
    host_dirty: bool,
    device_dirty: bool,
}

/// `recs_gpu_config` with the reference's constants filled in (`recs_gpu_default_config`).
#[derive(Debug, Clone, Copy)]
pub struct GpuConfig(pub sys::recs_gpu_config);

impl Default for GpuConfig {
    fn default() -> Self {
        let mut cfg = std::mem::MaybeUninit::<sys::recs_gpu_config>::uninit();
        // SAFETY: recs_gpu_default_config writes every field.
        unsafe {
            sys::recs_gpu_default_config(cfg.as_mut_ptr());
            GpuConfig(cfg.assume_init())
        }
    }
}

/// A failed call, with the library's sticky error text.
#[derive(thiserror::Error, Debug)]
#[error("recs_gpu status {status}: {message}")]
pub struct GpuError {
    pub status: i32,
    pub message: String,
}

impl GpuContext {
    pub fn new(config: GpuConfig) -> Result<Arc<Self>, GpuError> {
        let mut raw = std::ptr::null_mut();
        // SAFETY: plain C call; `raw` is written on success only.
        let status = unsafe { sys::recs_gpu_create(&config.0, &mut raw) };
        if status != sys::RECS_OK {
            return Err(GpuError {
                status,
                message: "no usable sm_100a device (there is no CPU fallback)".to_owned(),
            });
        }
        Ok(Arc::new(Self {
            raw,
            residency: Mutex::new(HashMap::new()),
        }))
    }

    fn check(&self, status: i32) -> Result<(), GpuError> {
        if status == sys::RECS_OK {
            return Ok(());
        }
        // SAFETY: never NULL, owned by the context.
        let message = unsafe { CStr::from_ptr(sys::recs_gpu_last_error(self.raw)) }
            .to_string_lossy()
            .into_owned();
        Err(GpuError { status, message })
    }

    /// Residency of one column before a GPU system runs over it: (re)bind when the host Vec moved or changed length,
    /// upload when the host holds the newest bytes.  The caller holds the column's lock for the duration of the call
    /// (`World::borrow_component_vecs(_mut)`, crates/ecs/src/lib.rs:807-833).
    fn ensure_on_device(&self, archetype: u32, component: u64, column: RawColumn, will_write: bool) -> Result<(), GpuError> {
        let mut residency = self.residency.lock().expect("lock should not be poisoned");
        let entry = residency.entry((archetype, component)).or_insert(Residency {
            host_ptr: 0,
            len: u64::MAX,
            host_dirty: true,
            device_dirty: false,
        });
        if entry.host_ptr != column.ptr as usize || entry.len != column.len {
            // SAFETY: ptr/len describe a live Vec<T> the caller has locked.
            self.check(unsafe {
                sys::recs_gpu_bind_column(self.raw, archetype, component, column.ptr, column.elem_size, column.len)
            })?;
            *entry = Residency {
                host_ptr: column.ptr as usize,
                len: column.len,
                host_dirty: true,
                device_dirty: false,
            };
        }
        if entry.host_dirty {
            self.check(unsafe { sys::recs_gpu_upload(self.raw, archetype, component) })?;
            entry.host_dirty = false;
        }
        entry.device_dirty |= will_write;
        Ok(())
    }

    /// Before a CPU system (or the user) reads a column a GPU system wrote: `recs_gpu_download`.
    pub fn ensure_on_host(&self, archetype: u32, component: u64) -> Result<(), GpuError> {
        let mut residency = self.residency.lock().expect("lock should not be poisoned");
        if let Some(entry) = residency.get_mut(&(archetype, component)) {
            if entry.device_dirty {
                self.check(unsafe { sys::recs_gpu_download(self.raw, archetype, component) })?;
                entry.device_dirty = false;
            }
        }
        Ok(())
    }

    /// A CPU system is about to write the column: the device copy goes stale.
    pub fn mark_host_dirty(&self, archetype: u32, component: u64) {
        if let Some(entry) = self.residency.lock().expect("lock should not be poisoned").get_mut(&(archetype, component)) {
            entry.host_dirty = true;
        }
    }

    /// Command playback (crates/ecs/src/systems/command_buffers.rs:377-435) happened on the host World: replay the
    /// structure on the resident columns without a round trip -- appended rows go up alone, the batch of
    /// `swap_remove`s (crates/ecs/src/archetypes.rs:233-243) is replayed as row moves.
    pub fn mirror_playback(
        &self,
        archetype: u32,
        columns: &[(u64, RawColumn)],
        moves: &[(u32, u32)], // (dst row, src row), sources as they were before the batch
        new_len: u64,
    ) -> Result<(), GpuError> {
        let mut residency = self.residency.lock().expect("lock should not be poisoned");
        let mut resident = Vec::new();
        for (component, column) in columns {
            let Some(entry) = residency.get_mut(&(archetype, *component)) else { continue };
            if entry.host_dirty {
                continue; // the host copy is the newer one: it is re-bound and uploaded on next use anyway
            }
            if column.len > entry.len {
                self.check(unsafe { sys::recs_gpu_column_append(self.raw, archetype, *component, column.ptr, column.len) })?;
            } else {
                self.check(unsafe { sys::recs_gpu_column_rebind_host(self.raw, archetype, *component, column.ptr) })?;
            }
            entry.host_ptr = column.ptr as usize;
            entry.len = new_len;
            resident.push(*component);
        }
        if !moves.is_empty() && !resident.is_empty() {
            let dst: Vec<u32> = moves.iter().map(|m| m.0).collect();
            let src: Vec<u32> = moves.iter().map(|m| m.1).collect();
            self.check(unsafe {
                sys::recs_gpu_archetype_move_rows(
                    self.raw,
                    archetype,
                    resident.as_ptr(),
                    resident.len() as u32,
                    dst.as_ptr(),
                    src.as_ptr(),
                    moves.len() as u32,
                    new_len,
                )
            })?;
        }
        Ok(())
    }

    pub fn synchronize(&self) -> Result<(), GpuError> {
        self.check(unsafe { sys::recs_gpu_synchronize(self.raw) })
    }
}

impl Drop for GpuContext {
    fn drop(&mut self) {
        // `Drop for WorkerPool` joins its threads (crates/scheduler/src/executor/mod.rs:84-99); this waits for the stream.
        unsafe { sys::recs_gpu_destroy(self.raw) };
    }
}

/// A component column as the C ABI sees it: `Vec<T>` base pointer, `size_of::<T>()`, length.
#[derive(Debug, Clone, Copy)]
pub struct RawColumn {
    pub ptr: *mut c_void,
    pub elem_size: u32,
    pub len: u64,
}

/// What crates/ecs has to expose for this file (three thin accessors next to `World::borrow_component_vecs`):
pub trait WorldColumns {
    /// Base pointer / element size / length of one component column of one archetype, taken under the column's
    /// `RwLock` write guard (crates/ecs/src/archetypes.rs:15, 106-138), or `None` if the archetype lacks the component.
    fn raw_column(&self, archetype: ArchetypeIndex, component: TypeId) -> Option<RawColumn>;
    /// `archetype.entities` as `generation << 32 | id` (what `Entity::hash` feeds, crates/ecs/src/lib.rs:918-927).
    fn entity_keys(&self, archetype: ArchetypeIndex) -> Vec<u64>;
    /// `Entity` of a component index of an archetype (for playing back GPU-decided removals).
    fn entity_at(&self, archetype: ArchetypeIndex, row: u32) -> Option<Entity>;
}

/// A stable u64 per component type (the C ABI's `component_id`).
pub fn component_id<T: 'static>() -> u64 {
    use std::hash::{Hash, Hasher};
    let mut hasher = std::collections::hash_map::DefaultHasher::new();
    TypeId::of::<T>().hash(&mut hasher);
    hasher.finish() & !(1 << 63) // never collides with RECS_GPU_ENTITY_COMPONENT
}

// ---------------------------------------------------------------------------------------------------------------------
// impl System for GpuSystem
// ---------------------------------------------------------------------------------------------------------------------

/// The built-in kinds (`recs_gpu_system_kind`).
#[derive(Debug, Clone, Copy, PartialEq, Eq)]
pub enum GpuKind {
    NBodyGravity,
    NBodyMovement,
    NBodyAcceleration,
    RainGravity,
    RainWind,
    RainMovement,
    RainWindDirection,
    QueryAdd,
    NBodyCollision,
    RainGroundCollision,
}

impl GpuKind {
    fn raw(self) -> i32 {
        match self {
            GpuKind::NBodyGravity => sys::RECS_SYS_NBODY_GRAVITY,
            GpuKind::NBodyMovement => sys::RECS_SYS_NBODY_MOVEMENT,
            GpuKind::NBodyAcceleration => sys::RECS_SYS_NBODY_ACCELERATION,
            GpuKind::RainGravity => sys::RECS_SYS_RAIN_GRAVITY,
            GpuKind::RainWind => sys::RECS_SYS_RAIN_WIND,
            GpuKind::RainMovement => sys::RECS_SYS_RAIN_MOVEMENT,
            GpuKind::RainWindDirection => sys::RECS_SYS_RAIN_WIND_DIRECTION,
            GpuKind::QueryAdd => sys::RECS_SYS_QUERY_ADD,
            GpuKind::NBodyCollision => sys::RECS_SYS_NBODY_COLLISION,
            GpuKind::RainGroundCollision => sys::RECS_SYS_RAIN_GROUND_COLLISION,
        }
    }
}

/// One component a GPU system touches: the Rust type behind the C ABI's `component_id`.
#[derive(Debug, Clone)]
pub struct GpuComponent {
    pub id: u64,
    pub type_id: TypeId,
    pub type_name: &'static str,
}

impl GpuComponent {
    pub fn of<T: 'static>() -> Self {
        Self {
            id: component_id::<T>(),
            type_id: TypeId::of::<T>(),
            type_name: std::any::type_name::<T>(),
        }
    }

    /// Placeholder for a `Commands` parameter of a generic system (RECS_KPARAM_COMMANDS carries no component; its
    /// element size / type name describe the record `commands.create(record)` takes, or 0 / "" for remove() only).
    pub fn commands() -> Self {
        Self {
            id: 0,
            type_id: TypeId::of::<()>(),
            type_name: "Commands",
        }
    }
}

/// A system whose body runs on the GPU: replaces e.g. `n_body::gravity` (crates/n-body/src/lib.rs:104-135) one for one.
#[derive(Debug, Clone)]
pub struct GpuSystem {
    ctx: Arc<GpuContext>,
    id: u32,
    name: String,
    /// the components behind the ids `recs_gpu_system_describe` reports, in parameter order
    components: Vec<GpuComponent>,
    accesses: Vec<ComponentAccessDescriptor>,
    writes: Vec<bool>,
    /// `<Parameters as SystemParameters>::get_archetype_indices` of the CPU system this one replaces
    /// (crates/ecs/src/systems/mod.rs:986-1000): membership stays with the unchanged Rust code, filters included.
    outer: fn(&World) -> Vec<ArchetypeIndex>,
    /// the nested `Query<..>` parameter's own match (systems/mod.rs:1042), for gravity and collision
    nested: Option<fn(&World) -> Vec<ArchetypeIndex>>,
    /// completion must be host-visible (a CPU system depends on this one), or stream order suffices
    sync: bool,
    command_buffer: CommandBuffer,
    command_receiver: CommandReceiver,
}

impl GpuSystem {
    /// Registers a built-in kind.  `components` in the order include/recs_gpu.h lists next to the kind, e.g.
    /// `[Position, Acceleration, Mass]` for gravity.
    pub fn new(
        ctx: &Arc<GpuContext>,
        kind: GpuKind,
        components: Vec<GpuComponent>,
        outer: fn(&World) -> Vec<ArchetypeIndex>,
        nested: Option<fn(&World) -> Vec<ArchetypeIndex>>,
    ) -> Result<Self, GpuError> {
        let ids: Vec<u64> = components.iter().map(|c| c.id).collect();
        let mut id = 0u32;
        ctx.check(unsafe { sys::recs_gpu_system_create(ctx.raw, kind.raw(), ids.as_ptr(), ids.len() as u32, &mut id) })?;
        Self::from_registered(ctx, id, components, outer, nested)
    }

    /// Registers a generic system: the body as CUDA C++ source, compiled with NVRTC (`recs_gpu_system_create_kernel`).
    /// `params`: (component, kind = RECS_KPARAM_*, C++ type name in `source`).
    pub fn from_source(
        ctx: &Arc<GpuContext>,
        name: &str,
        source: &str,
        body: &str,
        params: &[(GpuComponent, u32, u32, &str)], // component, size_of::<T>(), kind, C++ type
        outer: fn(&World) -> Vec<ArchetypeIndex>,
    ) -> Result<Self, GpuError> {
        let type_names: Vec<CString> = params.iter().map(|p| CString::new(p.3).expect("no NUL")).collect();
        let raw: Vec<sys::recs_gpu_kernel_param> = params
            .iter()
            .zip(&type_names)
            .map(|(p, type_name)| sys::recs_gpu_kernel_param {
                component_id: p.0.id,
                elem_size: p.1,
                kind: p.2,
                type_name: type_name.as_ptr(),
            })
            .collect();
        let (name_c, source_c, body_c) = (
            CString::new(name).expect("no NUL"),
            CString::new(source).expect("no NUL"),
            CString::new(body).expect("no NUL"),
        );
        let mut id = 0u32;
        ctx.check(unsafe {
            sys::recs_gpu_system_create_kernel(
                ctx.raw,
                name_c.as_ptr(),
                source_c.as_ptr(),
                body_c.as_ptr(),
                raw.as_ptr(),
                raw.len() as u32,
                std::ptr::null(),
                0,
                &mut id,
            )
        })?;
        let components = params
            .iter()
            .filter(|p| p.2 == sys::RECS_KPARAM_READ || p.2 == sys::RECS_KPARAM_WRITE)
            .map(|p| p.0.clone())
            .collect();
        Self::from_registered(ctx, id, components, outer, None)
    }

    fn from_registered(
        ctx: &Arc<GpuContext>,
        id: u32,
        components: Vec<GpuComponent>,
        outer: fn(&World) -> Vec<ArchetypeIndex>,
        nested: Option<fn(&World) -> Vec<ArchetypeIndex>>,
    ) -> Result<Self, GpuError> {
        // `System::name()` / `System::component_accesses()` come back from the library (parameter order,
        // crates/ecs/src/systems/mod.rs:973-978), so the schedule sees exactly what the kernels touch.
        let mut desc = std::mem::MaybeUninit::<sys::recs_gpu_system_desc>::zeroed();
        ctx.check(unsafe { sys::recs_gpu_system_describe(ctx.raw, id, desc.as_mut_ptr()) })?;
        let desc = unsafe { desc.assume_init() };
        let name = unsafe { CStr::from_ptr(desc.name) }.to_string_lossy().into_owned();
        let mut accesses = Vec::new();
        let mut writes = Vec::new();
        let mut ordered = Vec::new();
        for access in &desc.access[..desc.n_access as usize] {
            let component = components
                .iter()
                .find(|c| c.id == access.component_id)
                .expect("every accessed component id was passed at registration")
                .clone();
            accesses.push(if access.is_write != 0 {
                ComponentAccessDescriptor::Write(component.type_id, component.type_name.to_owned())
            } else {
                ComponentAccessDescriptor::Read(component.type_id, component.type_name.to_owned())
            });
            writes.push(access.is_write != 0);
            ordered.push(component);
        }
        let (command_buffer, command_receiver) = ecs::systems::command_buffers::command_channel();
        Ok(Self {
            ctx: Arc::clone(ctx),
            id,
            name,
            components: ordered,
            accesses,
            writes,
            outer,
            nested,
            sync: true,
            command_buffer,
            command_receiver,
        })
    }

    /// Every dependent of this system is a GPU system of the same context: stream order stands in for the drop of
    /// `SystemExecutionGuard::finished_sender` (crates/ecs/src/lib.rs:525-557) and `run` returns after the enqueue.
    pub fn asynchronous(mut self) -> Self {
        self.sync = false;
        self
    }
}

impl System for GpuSystem {
    fn name(&self) -> &str {
        &self.name
    }

    fn component_accesses(&self) -> Vec<ComponentAccessDescriptor> {
        self.accesses.clone()
    }

    fn try_as_sequentially_iterable(&self) -> Option<Box<dyn SequentiallyIterable>> {
        Some(Box::new(self.clone()))
    }

    fn try_as_segment_iterable(&self) -> Option<Box<dyn SegmentIterable>> {
        None // one task per tick: the device does the fan-out (executor/mod.rs:276-284 takes the sequential path)
    }

    fn command_buffer(&self) -> CommandBuffer {
        self.command_buffer.clone()
    }

    fn command_receiver(&self) -> CommandReceiver {
        self.command_receiver.clone()
    }
}

impl SequentiallyIterable for GpuSystem {
    fn run(&self, world: &World) -> SystemResult<()> {
        let outer = (self.outer)(world);
        let nested = self.nested.map(|query| query(world));
        // 1. residency of every column the system touches, for every matched archetype
        for (archetypes, range) in [(Some(&outer), 0..self.outer_access_count()), (nested.as_ref(), self.outer_access_count()..self.components.len())] {
            let Some(archetypes) = archetypes else { continue };
            for &archetype in archetypes {
                for slot in range.clone() {
                    let component = &self.components[slot];
                    let column = world
                        .raw_column(archetype, component.type_id)
                        .ok_or(SystemError::BorrowedData)?;
                    self.ctx
                        .ensure_on_device(archetype as u32, component.id, column, self.writes[slot])
                        .map_err(to_system_error)?;
                }
            }
        }
        // 2. membership, computed by the unchanged Rust code; ArchetypeIndex (usize) narrows to u32: the reference
        //    allocates archetype indices densely from 0 (crates/ecs/src/archetypes.rs:520-543)
        let outer32: Vec<u32> = outer.iter().map(|&a| a as u32).collect();
        self.ctx
            .check(unsafe { sys::recs_gpu_system_set_archetypes(self.ctx.raw, self.id, outer32.as_ptr(), outer32.len() as u32) })
            .map_err(to_system_error)?;
        if let Some(nested) = nested {
            let nested32: Vec<u32> = nested.iter().map(|&a| a as u32).collect();
            self.ctx
                .check(unsafe {
                    sys::recs_gpu_system_set_query_archetypes(self.ctx.raw, self.id, nested32.as_ptr(), nested32.len() as u32)
                })
                .map_err(to_system_error)?;
        }
        // 3. launch on the context stream
        let flags = if self.sync { sys::RECS_GPU_SYNC } else { sys::RECS_GPU_ASYNC };
        self.ctx
            .check(unsafe { sys::recs_gpu_run_system(self.ctx.raw, self.id, flags) })
            .map_err(to_system_error)?;
        // 4. `Commands::remove` decided on the device (collision, ground_collision): fetch the rows, record the
        //    removals in this system's own command buffer; playback is the reference's (lib.rs:321-323)
        if self.decides_removals() {
            for &archetype in &outer {
                let mut count = 0u32;
                self.ctx
                    .check(unsafe {
                        sys::recs_gpu_system_removals(self.ctx.raw, self.id, archetype as u32, std::ptr::null_mut(), 0, &mut count)
                    })
                    .map_err(to_system_error)?;
                let mut rows = vec![0u32; count as usize];
                self.ctx
                    .check(unsafe {
                        sys::recs_gpu_system_removals(self.ctx.raw, self.id, archetype as u32, rows.as_mut_ptr(), count, &mut count)
                    })
                    .map_err(to_system_error)?;
                for row in rows {
                    if let Some(entity) = world.entity_at(archetype, row) {
                        self.command_buffer.remove(entity);
                    }
                }
            }
        }
        Ok(())
    }
}

impl GpuSystem {
    /// Accesses of the outer parameters come first, the nested query's after them (systems/mod.rs:973-978).
    fn outer_access_count(&self) -> usize {
        match self.nested {
            None => self.components.len(),
            Some(_) if self.name == "gravity" => 2,   // Read<Position>, Write<Acceleration> | Position, Mass
            Some(_) => 1,                             // collision: Read<Position> | Position
        }
    }

    fn decides_removals(&self) -> bool {
        self.name == "collision" || self.name == "ground_collision"
    }
}

fn to_system_error(error: GpuError) -> SystemError {
    match error.status {
        // RECS_ERR_MISSING_PARAMETER mirrors SystemError::MissingParameter; everything else surfaces like a dead
        // worker (crates/scheduler/src/executor/mod.rs:223-233)
        sys::RECS_ERR_MISSING_PARAMETER | sys::RECS_ERR_STALE_BINDING | sys::RECS_ERR_SIZE_MISMATCH => SystemError::BorrowedData,
        _ => SystemError::Panic,
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// with_gpu(): builder add-on and application decorator (the shape of crates/gfx-plugin/src/lib.rs:44-98, 214-377)
// ---------------------------------------------------------------------------------------------------------------------

/// Extension of any `ApplicationBuilder`, like `Graphical::with_rendering` (crates/gfx-plugin/src/lib.rs:44-60).
pub trait Gpu: ApplicationBuilder + Sized {
    fn with_gpu(self, config: GpuConfig) -> Result<GpuApplicationBuilder<Self>, GpuError> {
        Ok(GpuApplicationBuilder {
            app_builder: self,
            ctx: GpuContext::new(config)?,
            systems: Vec::new(),
        })
    }
}

impl<Builder: ApplicationBuilder> Gpu for Builder {}

#[derive(Debug)]
pub struct GpuApplicationBuilder<AppBuilder> {
    app_builder: AppBuilder,
    ctx: Arc<GpuContext>,
    systems: Vec<GpuSystem>,
}

impl<AppBuilder: ApplicationBuilder> GpuApplicationBuilder<AppBuilder> {
    /// Registers a built-in GPU system in place of the CPU system `Replaced` (its parameter tuple gives the archetype
    /// match).  Example, the benchmark's registration (crates/ecs_bench_suite/src/recs/n_body.rs:54-59):
    ///
    /// ```ignore
    /// BasicApplicationBuilder::default()
    ///     .with_gpu(GpuConfig::default())?
    ///     .add_gpu_system::<(Write<Position>, Read<Velocity>)>(GpuKind::NBodyMovement,
    ///         vec![GpuComponent::of::<Position>(), GpuComponent::of::<Velocity>()], None)?
    ///     .add_gpu_system::<(Write<Velocity>, Read<Acceleration>)>(GpuKind::NBodyAcceleration,
    ///         vec![GpuComponent::of::<Velocity>(), GpuComponent::of::<Acceleration>()], None)?
    ///     .add_gpu_system::<(Read<Position>, Write<Acceleration>)>(GpuKind::NBodyGravity,
    ///         vec![GpuComponent::of::<Position>(), GpuComponent::of::<Acceleration>(), GpuComponent::of::<Mass>()],
    ///         Some(<(Read<Position>, Read<Mass>) as SystemParameters>::get_archetype_indices))?
    ///     .add_system(benchmark_system)
    ///     .build();
    /// ```
    pub fn add_gpu_system<Replaced: SystemParameters>(
        mut self,
        kind: GpuKind,
        components: Vec<GpuComponent>,
        nested: Option<fn(&World) -> Vec<ArchetypeIndex>>,
    ) -> Result<Self, GpuError> {
        let system = GpuSystem::new(&self.ctx, kind, components, Replaced::get_archetype_indices, nested)?;
        self.systems.push(system);
        Ok(self)
    }

    /// Ordinary CPU systems pass through to the wrapped builder.
    pub fn add_system<S, Parameters>(mut self, system: S) -> Self
    where
        S: IntoSystem<Parameters>,
        Parameters: SystemParameters,
    {
        self.app_builder = self.app_builder.add_system(system);
        self
    }

    pub fn build(self) -> GpuApplication<AppBuilder::App> {
        GpuApplication {
            application: self.app_builder.build(),
            ctx: self.ctx,
            systems: self.systems,
        }
    }
}

/// Decorator of any `Application`: delegates everything, injects the GPU systems before `run`
/// (as `GraphicalApplication` does with its rendering systems, crates/gfx-plugin/src/lib.rs:373-376).
#[derive(Debug)]
pub struct GpuApplication<App> {
    application: App,
    ctx: Arc<GpuContext>,
    systems: Vec<GpuSystem>,
}

impl<App: Application> GpuApplication<App> {
    pub fn context(&self) -> &Arc<GpuContext> {
        &self.ctx
    }
}

impl<App> Application for GpuApplication<App>
where
    App: Application + AddBoxedSystem,
{
    type Error = App::Error;

    fn create_empty_entity(&mut self) -> Result<Entity, Self::Error> {
        self.application.create_empty_entity()
    }

    fn create_entity(&mut self, components: impl IntoBoxedComponentIter) -> Result<Entity, Self::Error> {
        self.application.create_entity(components)
    }

    fn create_entities(
        &mut self,
        creations: impl IntoIterator<Item = impl IntoBoxedComponentIter>,
    ) -> Result<Vec<Entity>, Self::Error> {
        self.application.create_entities(creations)
    }

    fn remove_entities(&mut self, entities: impl IntoIterator<Item = Entity>) -> Result<(), Self::Error> {
        self.application.remove_entities(entities)
    }

    fn add_component<ComponentType: Debug + Send + Sync + 'static>(
        &mut self,
        entity: Entity,
        component: ComponentType,
    ) -> Result<(), Self::Error> {
        // the entity changes archetype (crates/ecs/src/archetypes.rs:474-518): rows of two archetypes move; the next
        // `GpuSystem::run` sees the new Vec pointers / lengths in `ensure_on_device` and re-binds those columns
        self.application.add_component(entity, component)
    }

    fn remove_component<ComponentType: Debug + Send + Sync + 'static>(&mut self, entity: Entity) -> Result<(), Self::Error> {
        self.application.remove_component::<ComponentType>(entity)
    }

    fn add_system<S, Parameters>(&mut self, system: S)
    where
        S: IntoSystem<Parameters>,
        Parameters: SystemParameters,
    {
        self.application.add_system(system);
    }

    fn run<E: Executor + 'static, S: Schedule + 'static>(mut self, shutdown_receiver: Receiver<()>) -> Result<(), Self::Error> {
        // GPU systems are `Box<dyn System>` like any other: `S::generate` orders them with the CPU systems
        for system in self.systems.drain(..) {
            self.application.add_boxed_system(Box::new(system));
        }
        let result = self.application.run::<E, S>(shutdown_receiver);
        let _ = self.ctx.synchronize();
        result
    }
}

/// `BasicApplication::systems` is `pub` (crates/ecs/src/lib.rs:247): pushing an already boxed system is one line.
pub trait AddBoxedSystem {
    fn add_boxed_system(&mut self, system: Box<dyn System>);
}

impl AddBoxedSystem for ecs::BasicApplication {
    fn add_boxed_system(&mut self, system: Box<dyn System>) {
        self.systems.push(system);
    }
}

// Accessors to be added inside crates/ecs (World's fields are crate-private); shown against the reference's storage:
//
// impl WorldColumns for World {
//     fn raw_column(&self, archetype: ArchetypeIndex, component: TypeId) -> Option<RawColumn> {
//         let archetype = self.archetypes.get(archetype)?;
//         let column = archetype.component_typeid_to_component_vec.get(&component)?;      // archetypes.rs:178-182
//         Some(column.as_raw())   // ComponentVec::as_raw: (Vec::as_mut_ptr, size_of::<T>(), Vec::len) under try_write()
//     }
//     fn entity_keys(&self, archetype: ArchetypeIndex) -> Vec<u64> {
//         self.archetypes[archetype].entities.iter().map(|e| (e.generation as u64) << 32 | e.id as u64).collect()
//     }
//     fn entity_at(&self, archetype: ArchetypeIndex, row: u32) -> Option<Entity> {
//         self.archetypes.get(archetype)?.entities.get(row as usize).copied()
//     }
// }
use WorldColumns as _;
