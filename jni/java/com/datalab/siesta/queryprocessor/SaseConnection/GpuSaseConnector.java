package com.datalab.siesta.queryprocessor.SaseConnection;

import com.datalab.siesta.queryprocessor.model.Constraints.Constraint;
import com.datalab.siesta.queryprocessor.model.Constraints.GapConstraint;
import com.datalab.siesta.queryprocessor.model.Constraints.TimeConstraint;
import com.datalab.siesta.queryprocessor.model.Events.Event;
import com.datalab.siesta.queryprocessor.model.Events.EventBoth;
import com.datalab.siesta.queryprocessor.model.Events.EventPos;
import com.datalab.siesta.queryprocessor.model.Events.EventSymbol;
import com.datalab.siesta.queryprocessor.model.Events.EventTs;
import com.datalab.siesta.queryprocessor.model.Occurrence;
import com.datalab.siesta.queryprocessor.model.Occurrences;
import com.datalab.siesta.queryprocessor.model.Patterns.ComplexPattern;
import com.datalab.siesta.queryprocessor.model.Patterns.SIESTAPattern;
import com.datalab.siesta.queryprocessor.model.Patterns.SimplePattern;
import com.datalab.siesta.queryprocessor.model.Utils.Utils;
import org.springframework.beans.factory.annotation.Autowired;
import org.springframework.beans.factory.config.ConfigurableBeanFactory;
import org.springframework.context.annotation.Primary;
import org.springframework.context.annotation.Scope;
import org.springframework.stereotype.Service;

import java.sql.Timestamp;
import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Locale;
import java.util.Map;

/**
 * Drop-in for the verification seam: same bean type, same method, same result.
 *
 *   List&lt;Occurrences&gt; evaluate(SIESTAPattern pattern, Map&lt;String, List&lt;Event&gt;&gt; events, boolean onlyAppearances)
 *
 * (SaseConnector.java:48).  @Primary makes Spring inject it wherever a SaseConnector is asked for, so
 * QueryPlanPatternDetection (:89, :121, :235), ...Single, ...Groups and QueryPlanExplorationAccurate run unchanged.
 *
 * What differs from the CPU path, and why the callers cannot tell:
 *  - the library evaluates AND selects (Occurrences.clearOccurrences) in one pass.  The seam does not know
 *    `returnAll`, so the request always runs with SIESTA_F_RETURN_ALL: the occurrences returned are
 *    clearOccurrences(true)'s selection, whose first element is clearOccurrences(false)'s.  Every caller applies
 *    clearOccurrences(returnAll) to what it gets (QueryPlanPatternDetection.java:122, ...Single.java:45-46,
 *    ...Groups.java:51-52) - on this list that is the identity (returnAll = true) or "keep the first" (false) - or
 *    only reads the trace ids (retrieveTimeInformation :235-236).
 *  - the events travel as a CSR log: trace offsets, dense activity ids (names folded with toLowerCase, because the
 *    engine compares types with equalsIgnoreCase, State.java:135-137), epoch milliseconds.
 *  - a trace beyond the GPU engine's per-trace limits does not fail the request: the library answers all others and
 *    lists it; this class hands the listed traces to super.evaluate (the reference's engine).
 *  - EventPos lists (positions-mode logs) carry their own `position`; the CSR has no position column - the position
 *    of an event is its index in the trace - so such a list is laid out sparsely: slot = position, empty slots hold an
 *    activity id no pattern uses.  The library's EventPos route (SIESTA_F_EVT_POS) then sees exactly the
 *    (position, index-in-list) pairs Utils.transformToSaseEvents produces (Utils.java:59-62).
 *
 * NOT compiled in the repository this file ships in (no JDK in its image); the native half is compiled there.
 */
@Service
@Primary
@Scope(value = ConfigurableBeanFactory.SCOPE_PROTOTYPE)
public class GpuSaseConnector extends SaseConnector {

    /** one native context for the JVM: every GPU of the box */
    private static final long MULTI = GpuNative.init(visibleDevices());

    @Autowired
    public GpuSaseConnector(Utils utils) {
        super(utils);
    }

    private static int[] visibleDevices() {
        int n = Integer.parseInt(System.getProperty("siesta.gpu.devices", "1"));
        int[] ids = new int[n];
        for (int i = 0; i < n; i++) ids[i] = i;
        return ids;
    }

    @Override
    public List<Occurrences> evaluate(SIESTAPattern pattern, Map<String, List<Event>> events, boolean onlyAppearances) {
        return evaluate(pattern, events, onlyAppearances, true);
    }

    /**
     * The same request with the caller's returnAll known.  The seam's signature does not carry it, so the override above asks
     * for every non-overlapping occurrence; a plan that is adapted to call this overload (one line in
     * QueryPlanPatternDetection.execute :121, which reads qpdw.isReturnAll() on the next line anyway) gets
     * clearOccurrences(returnAll)'s selection directly, and with returnAll = false the library only computes the first-largest
     * occurrence of every trace (no relative seconds, no overlap test: the fastest kernels).
     */
    public List<Occurrences> evaluate(SIESTAPattern pattern, Map<String, List<Event>> events, boolean onlyAppearances, boolean returnAll) {
        // ---- dictionaries: trace ids in map order, activity names folded case-insensitively
        List<String> traceIds = new ArrayList<>();
        Map<String, Integer> actIds = new HashMap<>();
        List<String> actNames = new ArrayList<>();
        for (String name : pattern.getEventTypes()) intern(actIds, actNames, name);
        boolean evtPos = false;
        long nEvents = 0;
        for (Map.Entry<String, List<Event>> e : events.entrySet()) {
            if (e.getValue().isEmpty()) continue;               // SaseConnector.java:53-55
            traceIds.add(e.getKey());
            Event first = e.getValue().get(0);
            evtPos = !(first instanceof EventTs);               // Utils.transformToSaseEvents:51 / :59
            if (evtPos) {                                        // sparse layout: slot = position
                int last = ((EventPos) e.getValue().get(e.getValue().size() - 1)).getPosition();
                nEvents += last + 1;
            } else nEvents += e.getValue().size();
        }
        if (nEvents > Integer.MAX_VALUE - 8) throw new RuntimeException("request larger than a Java array: load the log resident (GpuNative.logLoad)");
        final int filler = actNames.size();                      // an activity no state of the pattern owns
        long[] traceOff = new long[traceIds.size() + 1];
        int[] act = new int[(int) nEvents];
        long[] tsMs = new long[(int) nEvents];
        int at = 0, t = 0;
        for (String id : traceIds) {
            List<Event> evs = events.get(id);
            if (evtPos) {
                int base = at, last = ((EventPos) evs.get(evs.size() - 1)).getPosition();
                java.util.Arrays.fill(act, base, base + last + 1, filler);
                for (Event ev : evs) act[base + ((EventPos) ev).getPosition()] = intern(actIds, actNames, ev.getName());
                at = base + last + 1;
            } else {
                for (Event ev : evs) {
                    act[at] = intern(actIds, actNames, ev.getName());
                    tsMs[at] = ((EventTs) ev).getTimestamp().getTime();
                    at++;
                }
            }
            traceOff[++t] = at;
        }
        // ---- the pattern: ComplexPattern.getNfa / SimplePattern.getNfa, compiled by the library
        int[] nfa = GpuNative.patternCompile(symbolsOf(pattern, actIds, actNames), constraintsOf(pattern), onlyAppearances);
        int flags = (returnAll ? GpuNative.F_RETURN_ALL : 0) | (evtPos ? GpuNative.F_EVT_POS : 0);
        long m = GpuNative.evaluateEvents(MULTI, traceOff, act, tsMs, actNames.size() + 1, nfa, flags);
        try {
            long[] sizes = GpuNative.matchesSizes(m);
            if (sizes[4] > 0)   // traces on which the Java engine throws: the reference fails the request (SaseConnector.java:60-62)
                throw new RuntimeException("SASE engine error on " + sizes[4] + " trace(s), first: " + traceIds.get((int) GpuNative.matchesLongs(m, 4)[0]));
            long[] tr = GpuNative.matchesLongs(m, 0), occOff = GpuNative.matchesLongs(m, 1), evOff = GpuNative.matchesLongs(m, 2);
            long[] ts = GpuNative.matchesLongs(m, 3);
            int[] pos = GpuNative.matchesInts(m, 0), rank = GpuNative.matchesInts(m, 1), a = GpuNative.matchesInts(m, 2);
            List<Occurrences> out = new ArrayList<>(tr.length);
            for (int i = 0; i < tr.length; i++) {
                String traceId = traceIds.get((int) tr[i]);
                Occurrences ocs = new Occurrences();
                ocs.setTraceID(traceId);
                for (long o = occOff[i]; o < occOff[i + 1]; o++) {
                    List<EventBoth> evs = new ArrayList<>();
                    for (long e = evOff[(int) o]; e < evOff[(int) o + 1]; e++) {
                        // SaseEvent.getEventBoth (SaseEvent.java:94-106): EventTs route -> timestamp * 1000 + minTs and
                        // position = index in the list; EventPos route -> position only
                        int k = (int) e;
                        evs.add(evtPos ? new EventBoth(actNames.get(a[k]), traceId, null, pos[k])
                                       : new EventBoth(actNames.get(a[k]), traceId, new Timestamp(ts[k]), rank[k]));
                    }
                    ocs.addOccurrence(new Occurrence(evs));
                }
                out.add(ocs);
            }
            // traces beyond the GPU engine's per-trace limits (more than 64 pattern-relevant events ...): the reference's own
            // engine answers them - it has no such limit (Engine.java:207-224).  Its result carries EVERY match, which the
            // caller's clearOccurrences reduces exactly as on the CPU path.
            if (sizes[5] > 0) {
                Map<String, List<Event>> rest = new HashMap<>();
                for (long u : GpuNative.matchesLongs(m, 5)) rest.put(traceIds.get((int) u), events.get(traceIds.get((int) u)));
                out.addAll(super.evaluate(pattern, rest, onlyAppearances));
            }
            return out;
        } finally {
            GpuNative.matchesFree(m);
        }
    }

    private static int intern(Map<String, Integer> ids, List<String> names, String name) {
        String k = name.toLowerCase(Locale.ROOT);
        Integer id = ids.get(k);
        if (id == null) {
            id = names.size();
            ids.put(k, id);
            names.add(name);
        }
        return id;
    }

    private static int[] symbolsOf(SIESTAPattern p, Map<String, Integer> ids, List<String> names) {
        List<EventSymbol> es;
        if (p instanceof ComplexPattern) es = ((ComplexPattern) p).getEventsWithSymbols();
        else {                                   // SimplePattern: every event is a "_" state (SimplePattern.java:96-104)
            es = new ArrayList<>();
            for (EventPos e : ((SimplePattern) p).getEvents()) es.add(new EventSymbol(e.getName(), e.getPosition(), "_"));
        }
        int[] out = new int[3 * es.size()];
        int i = 0;
        for (EventSymbol e : es) {
            out[i++] = intern(ids, names, e.getName());
            out[i++] = e.getPosition();
            String s = e.getSymbol();
            out[i++] = "+".equals(s) ? GpuNative.SYM_PLUS : "*".equals(s) ? GpuNative.SYM_STAR : "!".equals(s) ? GpuNative.SYM_NOT
                     : "||".equals(s) ? GpuNative.SYM_OR : GpuNative.SYM_NORMAL;
        }
        return out;
    }

    private static long[] constraintsOf(SIESTAPattern p) {
        List<Constraint> cs = p.getConstraints();
        long[] out = new long[6 * cs.size()];
        int i = 0;
        for (Constraint c : cs) {
            out[i++] = c.getPosA();
            out[i++] = c.getPosB();
            out[i++] = c instanceof TimeConstraint ? GpuNative.CONSTRAINT_TIME : GpuNative.CONSTRAINT_GAP;
            out[i++] = "within".equals(c.getMethod()) ? GpuNative.METHOD_WITHIN : GpuNative.METHOD_ATLEAST;   // SIESTAPattern.java:136-145
            out[i++] = c instanceof TimeConstraint ? ((TimeConstraint) c).getConstraint() : ((GapConstraint) c).getConstraint();
            String g = c instanceof TimeConstraint ? ((TimeConstraint) c).getGranularity() : "seconds";
            out[i++] = "minutes".equals(g) ? GpuNative.GRAN_MINUTES : "hours".equals(g) ? GpuNative.GRAN_HOURS : GpuNative.GRAN_SECONDS;
        }
        return out;
    }
}
