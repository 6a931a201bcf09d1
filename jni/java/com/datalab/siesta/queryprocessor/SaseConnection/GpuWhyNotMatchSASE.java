package com.datalab.siesta.queryprocessor.SaseConnection;

import com.datalab.siesta.queryprocessor.model.Constraints.Constraint;
import com.datalab.siesta.queryprocessor.model.Constraints.GapConstraint;
import com.datalab.siesta.queryprocessor.model.Constraints.TimeConstraint;
import com.datalab.siesta.queryprocessor.model.Events.Event;
import com.datalab.siesta.queryprocessor.model.Events.EventPos;
import com.datalab.siesta.queryprocessor.model.Events.EventTs;
import com.datalab.siesta.queryprocessor.model.Patterns.SimplePattern;
import com.datalab.siesta.queryprocessor.model.WhyNotMatch.AlmostMatch;
import com.datalab.siesta.queryprocessor.model.WhyNotMatch.UsingSase.UncertainTimeEvent;
import com.datalab.siesta.queryprocessor.model.WhyNotMatch.UsingSase.WhyNotMatchSASE;
import org.springframework.beans.factory.config.ConfigurableBeanFactory;
import org.springframework.context.annotation.Primary;
import org.springframework.context.annotation.Scope;
import org.springframework.stereotype.Service;

import java.util.ArrayList;
import java.util.HashMap;
import java.util.List;
import java.util.Locale;
import java.util.Map;

/**
 * Drop-in for the why-not-match seam: same bean type, same method, same result.
 *
 *   List&lt;AlmostMatch&gt; evaluate(SimplePattern sp, Map&lt;String, List&lt;Event&gt;&gt; restEvents, int uncertaintyPerEvent, int step, int k)
 *
 * (WhyNotMatchSASE.java:37).  QueryPlanWhyNotMatch.execute (:54-100) runs unchanged: it hands in the traces without a true
 * occurrence and puts what comes back into the response.
 *
 * The reference's engine turns every combination of uncertain events into a run object; the library computes the match it
 * would report (least total change, the last such match in emission order) with one sweep per start event (csrc/wnm.cuh).
 * Traces whose uncertain stream exceeds the library's bound go to super.evaluate.
 *
 * NOT compiled in the repository this file ships in (no JDK in its image); the native half is compiled there.
 */
@Service
@Primary
@Scope(value = ConfigurableBeanFactory.SCOPE_PROTOTYPE)
public class GpuWhyNotMatchSASE extends WhyNotMatchSASE {

    private static final long MULTI = GpuNative.init(new int[]{0});

    @Override
    public List<AlmostMatch> evaluate(SimplePattern sp, Map<String, List<Event>> restEvents, int uncertaintyPerEvent, int step, int k) {
        // constraints the library does not take (they do not name an earlier and a later event of the pattern): reference engine
        for (Constraint c : sp.getConstraints())
            if (c.getPosA() < 0 || c.getPosA() >= c.getPosB() || c.getPosB() >= sp.getEvents().size())
                return super.evaluate(sp, restEvents, uncertaintyPerEvent, step, k);
        Map<String, Integer> actIds = new HashMap<>();
        List<String> names = new ArrayList<>();
        int[] pattern = new int[sp.getEvents().size()];
        for (int i = 0; i < pattern.length; i++) pattern[i] = intern(actIds, names, sp.getEvents().get(i).getName());
        final int filler = names.size();
        List<String> traceIds = new ArrayList<>();
        boolean evtPos = false;
        long nEvents = 0;
        for (Map.Entry<String, List<Event>> e : restEvents.entrySet()) {
            if (e.getValue().isEmpty()) continue;
            traceIds.add(e.getKey());
            evtPos = !(e.getValue().get(0) instanceof EventTs);
            nEvents += evtPos ? ((EventPos) e.getValue().get(e.getValue().size() - 1)).getPosition() + 1 : e.getValue().size();
        }
        long[] traceOff = new long[traceIds.size() + 1];
        int[] act = new int[(int) nEvents];
        long[] tsMs = new long[(int) nEvents];
        int at = 0, t = 0;
        for (String id : traceIds) {
            List<Event> evs = restEvents.get(id);
            if (evtPos) {   // primary metric = position (EventPos.getPrimaryMetric): slot = position, empty slots hold no pattern activity
                int base = at, last = ((EventPos) evs.get(evs.size() - 1)).getPosition();
                java.util.Arrays.fill(act, base, base + last + 1, filler);
                for (Event ev : evs) act[base + ((EventPos) ev).getPosition()] = internOr(actIds, ev.getName(), filler);
                at = base + last + 1;
            } else {
                for (Event ev : evs) {
                    act[at] = internOr(actIds, ev.getName(), filler);
                    tsMs[at] = ((EventTs) ev).getTimestamp().getTime();
                    at++;
                }
            }
            traceOff[++t] = at;
        }
        long[] cons = new long[5 * sp.getConstraints().size()];
        int i = 0;
        for (Constraint c : sp.getConstraints()) {
            cons[i++] = c.getPosA();
            cons[i++] = c.getPosB();
            cons[i++] = c instanceof TimeConstraint ? 1 : 0;
            cons[i++] = "within".equals(c.getMethod()) ? 0 : 1;
            cons[i++] = c instanceof TimeConstraint ? ((TimeConstraint) c).getConstraintInSeconds() : ((GapConstraint) c).getConstraint();
        }
        long log = GpuNative.logLoad(MULTI, traceOff, act, tsMs, filler + 1);
        long a = 0;
        try {
            a = GpuNative.whyNotMatch(log, pattern, cons, uncertaintyPerEvent, step, k, null, evtPos ? GpuNative.F_EVT_POS : 0);
            long[] tr = GpuNative.almostLongs(a, 0), rest = GpuNative.almostLongs(a, 1);
            int[] value = GpuNative.almostInts(a, 2), change = GpuNative.almostInts(a, 3), spos = GpuNative.almostInts(a, 4);
            List<AlmostMatch> out = new ArrayList<>(tr.length);
            int m = pattern.length;
            for (int r = 0; r < tr.length; r++) {
                String traceId = traceIds.get((int) tr[r]);
                List<UncertainTimeEvent> evs = new ArrayList<>(m);
                for (int j = 0; j < m; j++)   // UncertainTimeEvent(trace_id, position, event_type, timestamp, isTimeStampSet, change), WhyNotMatchSASE.java:71-72
                    evs.add(new UncertainTimeEvent(traceId, spos[r * m + j], sp.getEvents().get(j).getName(), value[r * m + j], true, change[r * m + j]));
                out.add(new AlmostMatch(traceId, restEvents.get(traceId), evs));
            }
            if (rest.length > 0) {   // uncertain streams beyond the library's bound: the reference's own engine
                Map<String, List<Event>> more = new HashMap<>();
                for (long u : rest) more.put(traceIds.get((int) u), restEvents.get(traceIds.get((int) u)));
                out.addAll(super.evaluate(sp, more, uncertaintyPerEvent, step, k));
            }
            return out;
        } finally {
            if (a != 0) GpuNative.almostFree(a);
            GpuNative.logFree(log);
        }
    }

    private static int intern(Map<String, Integer> ids, List<String> names, String name) {
        String key = name.toLowerCase(Locale.ROOT);   // State.canStartWithEvent compares with equalsIgnoreCase (State.java:305)
        Integer id = ids.get(key);
        if (id == null) {
            id = names.size();
            ids.put(key, id);
            names.add(name);
        }
        return id;
    }

    private static int internOr(Map<String, Integer> ids, String name, int other) {
        Integer id = ids.get(name.toLowerCase(Locale.ROOT));
        return id == null ? other : id;
    }
}
