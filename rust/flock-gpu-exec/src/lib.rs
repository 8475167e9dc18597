//! flock-gpu-exec: run a Flock physical plan on a B200 behind DataFusion's `ExecutionPlan` surface.
//!
//! SOURCE ONLY: this image has neither `cargo` nor `rustc`, so this crate has never been compiled; the
//! tested contract is the C ABI (include/flockgpu.h, exercised through ctypes by tests/).  The shape
//! follows the only in-tree `impl ExecutionPlan` of the reference, `ShuffleWriterExec`
//! (playground/src/distributed_plan/shuffle_writer.rs:157-231), and the fork's `LambdaExecPlan::feed_batches`.
//!
//! `GpuPlanExec` wraps a whole plan (or sub-plan): its serde-JSON form -- the very string
//! `flock::runtime::context::marshal` produces (flock/src/runtime/context.rs:366-381) -- is handed to
//! `flock_context_unmarshal`; `execute()` exports the batches of its MemoryExec leaves through the Arrow C
//! Data Interface, calls the GPU executor and wraps the result in a one-batch stream.
//!
//! Two fallbacks keep the Lambda alive whatever the data looks like:
//!   * plan time: `rewrite_for_gpu` wraps the LARGEST sub-plans a parse-only `flock_context_unmarshal(NULL, json)`
//!     accepts; nodes the GPU path does not implement stay DataFusion nodes with GPU sub-plans below them;
//!   * run time: data-dependent refusals (a column with NULLs, LargeUtf8, a dictionary ...) surface as
//!     FLOCKGPU_ERR_UNSUPPORTED from feed / execute; `GpuPlanExec::execute` then runs the wrapped CPU plan instead.
pub mod ffi;
pub mod pinned;

use async_trait::async_trait;
use datafusion::arrow::array::{make_array_from_raw, ArrayRef, StructArray};
use datafusion::arrow::datatypes::SchemaRef;
use datafusion::arrow::ffi::{FFI_ArrowArray, FFI_ArrowSchema};
use datafusion::arrow::record_batch::RecordBatch;
use datafusion::error::{DataFusionError, Result};
use datafusion::physical_plan::memory::{MemoryExec, MemoryStream};
use datafusion::physical_plan::{DisplayFormatType, ExecutionPlan, Partitioning, SendableRecordBatchStream, Statistics};
use serde::{Deserialize, Serialize};
use std::any::Any;
use std::ffi::{CStr, CString};
use std::sync::{Arc, Mutex};

fn last_error() -> DataFusionError {
    let msg = unsafe { CStr::from_ptr(ffi::flockgpu_last_error()) }.to_string_lossy().into_owned();
    DataFusionError::Execution(msg) // maps to FlockError::Execution in context.rs:181
}

/// One GPU context per process (a Lambda instance handles one event at a time, cloud_context.rs:22).
/// The handle is an opaque pointer into the library, which serialises calls on a context with its own mutex: it may
/// cross threads (spawn_blocking) -- `Send` and `Copy` are asserted here, once, instead of at every use.
#[derive(Clone, Copy)]
struct Gpu(*mut ffi::flockgpu_ctx);
unsafe impl Send for Gpu {}
unsafe impl Sync for Gpu {}
static GPU: Mutex<Option<Gpu>> = Mutex::new(None);

fn gpu() -> Result<Gpu> {
    let mut g = GPU.lock().unwrap();
    if g.is_none() {
        let mut ctx = std::ptr::null_mut();
        if unsafe { ffi::flockgpu_open(0, &mut ctx) } != ffi::FLOCKGPU_OK {
            return Err(last_error());
        }
        *g = Some(Gpu(ctx));
    }
    Ok(*g.as_ref().unwrap())
}

/// A physical (sub-)plan executed on the GPU.
#[derive(Debug, Serialize, Deserialize)]
pub struct GpuPlanExec {
    /// The wrapped CPU plan: kept for schema/children/serde and as the fallback.
    plan: Arc<dyn ExecutionPlan>,
}

impl GpuPlanExec {
    /// Returns `None` when the GPU path does not support some node or expression of `plan`.
    pub fn try_new(plan: Arc<dyn ExecutionPlan>) -> Option<Self> {
        let json = CString::new(serde_json::to_string(&plan).ok()?).ok()?;
        let mut ec = std::ptr::null_mut();
        // parse-only unmarshal (ctx = NULL): validates that every node is supported without touching a device
        let rc = unsafe { ffi::flock_context_unmarshal(std::ptr::null_mut(), json.as_ptr(), &mut ec) };
        if rc != ffi::FLOCKGPU_OK {
            return None;
        }
        unsafe { ffi::flock_context_free(ec) };
        Some(Self { plan })
    }

    /// Leaves in breadth-first order, as `feed_data_sources` visits them (context.rs:262-266).
    fn leaves(&self) -> Vec<Arc<dyn ExecutionPlan>> {
        let mut out = vec![];
        let mut queue = std::collections::VecDeque::from(vec![self.plan.clone()]);
        while let Some(p) = queue.pop_front() {
            if p.children().is_empty() {
                out.push(p.clone());
            }
            queue.extend(p.children());
        }
        out
    }
}

#[async_trait]
#[typetag::serde(name = "gpu_plan_exec")]
impl ExecutionPlan for GpuPlanExec {
    fn as_any(&self) -> &dyn Any {
        self
    }
    fn as_mut_any(&mut self) -> &mut dyn Any {
        self
    }
    fn schema(&self) -> SchemaRef {
        self.plan.schema()
    }
    fn output_partitioning(&self) -> Partitioning {
        Partitioning::UnknownPartitioning(1) // one device partition
    }
    fn children(&self) -> Vec<Arc<dyn ExecutionPlan>> {
        // leaves stay reachable so that ExecutionContext::feed_data_sources finds the MemoryExec nodes
        self.plan.children()
    }
    fn with_new_children(&self, children: Vec<Arc<dyn ExecutionPlan>>) -> Result<Arc<dyn ExecutionPlan>> {
        Ok(Arc::new(Self { plan: self.plan.with_new_children(children)? }))
    }

    async fn execute(&self, partition: usize) -> Result<SendableRecordBatchStream> {
        assert_eq!(partition, 0);
        let gpu = gpu()?;
        let json = CString::new(serde_json::to_string(&self.plan).map_err(|e| DataFusionError::Execution(e.to_string()))?).unwrap();
        // 1. drain the MemoryExec leaves (they were filled by feed_data_sources)
        let mut per_leaf: Vec<Vec<RecordBatch>> = vec![];
        for leaf in self.leaves() {
            let mem = leaf.as_any().downcast_ref::<MemoryExec>().expect("leaf is a MemoryExec");
            per_leaf.push(mem.partitions().iter().flatten().cloned().collect());
        }
        let schema = self.schema();
        // 2. the blocking GPU call runs off the async executor
        let batches = tokio::task::spawn_blocking(move || -> Result<Option<Vec<RecordBatch>>> {
            let ctx = gpu.0;
            let mut ec = std::ptr::null_mut();
            if unsafe { ffi::flock_context_unmarshal(ctx, json.as_ptr(), &mut ec) } != 0 {
                return Err(last_error());
            }
            // export every batch through the C Data Interface (borrowed by the callee for the call only)
            let mut keep: Vec<(Vec<FFI_ArrowArray>, FFI_ArrowSchema)> = vec![];
            for batches in &per_leaf {
                if batches.is_empty() {
                    continue;
                }
                let s = FFI_ArrowSchema::try_from(batches[0].schema().as_ref()).map_err(DataFusionError::ArrowError)?;
                let arrays = batches.iter().map(|b| FFI_ArrowArray::new(StructArray::from(b.clone()).data())).collect();
                keep.push((arrays, s));
            }
            let schema_ptrs: Vec<*const FFI_ArrowSchema> = keep.iter().map(|(_, s)| s as *const _).collect();
            let array_ptrs: Vec<Vec<*const FFI_ArrowArray>> = keep.iter().map(|(a, _)| a.iter().map(|x| x as *const _).collect()).collect();
            let array_ptr_ptrs: Vec<*const *const FFI_ArrowArray> = array_ptrs.iter().map(|v| v.as_ptr()).collect();
            let counts: Vec<i32> = keep.iter().map(|(a, _)| a.len() as i32).collect();
            let mut out_table = std::ptr::null_mut();
            let rc = unsafe {
                let mut rc = ffi::flock_context_feed_data_sources(ec, schema_ptrs.as_ptr(), array_ptr_ptrs.as_ptr(), counts.as_ptr(), counts.len() as i32);
                if rc == 0 {
                    rc = ffi::flock_context_execute(ec, 0, &mut out_table);
                }
                rc
            };
            if rc == ffi::FLOCKGPU_ERR_UNSUPPORTED {
                // the DATA is outside what the GPU path implements (NULLs, LargeUtf8 ...): not an error, the CPU plan runs
                unsafe { ffi::flock_context_free(ec) };
                return Ok(None);
            }
            if rc != 0 {
                unsafe { ffi::flock_context_free(ec) };
                return Err(last_error());
            }
            // 3. import the result: the arrays are owned by the library until their release callback runs
            let mut out_array = FFI_ArrowArray::empty();
            let mut out_schema = FFI_ArrowSchema::empty();
            let rc = unsafe { ffi::flockgpu_table_export(ctx, out_table, 0, -1, &mut out_schema, &mut out_array) };
            unsafe {
                ffi::flockgpu_table_release(out_table);
                ffi::flock_context_clean_data_sources(ec);
                ffi::flock_context_free(ec);
            }
            if rc != 0 {
                return Err(last_error());
            }
            let array: ArrayRef = unsafe { make_array_from_raw(&out_array, &out_schema) }.map_err(DataFusionError::ArrowError)?;
            let s = array.as_any().downcast_ref::<StructArray>().expect("struct array");
            Ok(Some(vec![RecordBatch::from(s)]))
        })
        .await
        .map_err(|e| DataFusionError::Execution(e.to_string()))??;
        match batches {
            Some(batches) => Ok(Box::pin(MemoryStream::try_new(batches, schema, None)?)),
            // run-time fallback: the wrapped DataFusion plan, leaves already fed by feed_data_sources
            None => datafusion::physical_plan::collect(self.plan.clone()).await.and_then(|b| {
                Ok(Box::pin(MemoryStream::try_new(b, schema, None)?) as SendableRecordBatchStream)
            }),
        }
    }

    fn fmt_as(&self, _t: DisplayFormatType, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        write!(f, "GpuPlanExec: B200 (libflockgpu)")
    }
    fn statistics(&self) -> Statistics {
        Statistics::default()
    }
}

/// The fork's `LambdaExecPlan` (implemented in-tree by `ShuffleWriterExec`,
/// playground/src/distributed_plan/shuffle_writer.rs:157-161): a node whose input arrives as batches, not as a child.
impl datafusion::physical_plan::LambdaExecPlan for GpuPlanExec {
    fn feed_batches(&mut self, partitions: Vec<Vec<RecordBatch>>) {
        // hand the batches to the first MemoryExec leaf of the wrapped plan, exactly what feed_data_sources does for
        // a one-relation plan (flock/src/runtime/context.rs:293-303)
        if let Some(leaf) = self.leaves().into_iter().next() {
            unsafe {
                let mem = Arc::get_mut_unchecked(&mut leaf.clone());
                if let Some(m) = mem.as_mut_any().downcast_mut::<MemoryExec>() {
                    m.set_partitions(partitions);
                }
            }
        }
    }
}

/// Plan rewrite: called where Flock builds the per-function plan (flock/src/runtime/plan.rs:221-228).
/// Wraps the largest sub-plans the GPU path supports; everything else stays as DataFusion planned it.
pub fn rewrite_for_gpu(plan: Arc<dyn ExecutionPlan>) -> Arc<dyn ExecutionPlan> {
    if plan.children().is_empty() {
        return plan; // a bare MemoryExec gains nothing
    }
    if let Some(gpu) = GpuPlanExec::try_new(plan.clone()) {
        return Arc::new(gpu);
    }
    // this node (or something below it) is not implemented on the GPU: keep the node, rewrite its inputs
    let children: Vec<_> = plan.children().into_iter().map(rewrite_for_gpu).collect();
    plan.with_new_children(children).unwrap_or(plan)
}
